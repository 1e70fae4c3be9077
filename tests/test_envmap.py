"""EnvironmentLightColor::Texture (emitter.rs:300-427) + Distribution2D (math.rs:489-532): a lat-long image as the environment,
importance-sampled through its row-wise CDFs.

Layers as everywhere: (1) the oracle against the model (the spec atan2 / acos against numpy, pdf normalisation, samples follow the
pdf, a constant image == the constant environment), (2) the device arithmetic (tests/emu) == the oracle bit for bit, function by
function and on whole renders, (3) loaders.  The GPU render of the same scenes is in test_gpu.py."""
import json
import math
import os

import numpy as np
import pytest

import emu_binding as eb
from conftest import load_cbox, rel_l2
from oracle import binding as ob
from rustlight_b200 import SceneLoaderManager, _abi
from rustlight_b200.host import SceneError

STREAM = dict(estimator=ob.EST_STREAM, accel_mode=ob.ACCEL_BVH)


def sky_image(w=16, h=8, seed=0):
    """A small HDR sky: dim gradient, a bright 'sun' texel block, one black row (zero-probability conditionals) and one black texel."""
    rng = np.random.default_rng(seed)
    img = (0.05 + 0.3 * rng.random((h, w, 3))).astype(np.float32)
    img[1, 3:5] = [40.0, 35.0, 20.0]
    img[h - 2] = 0.0
    img[2, 7] = 0.0
    return img


def env_scene(img=None, w=32, h=32, area_light=False, floor=False):
    """An octahedron (diffuse) floating in the environment, optionally over a glossy floor and next to an area light."""
    P = [1, 0, 0, -1, 0, 0, 0, 1, 0, 0, -1, 0, 0, 0, 1, 0, 0, -1]
    idx = [0, 2, 4, 2, 1, 4, 1, 3, 4, 3, 0, 4, 2, 0, 5, 1, 2, 5, 3, 1, 5, 0, 3, 5]
    meshes = [{"material": {"type": "diffuse", "kd": [0.6, 0.5, 0.4]}, "indices": idx, "P": P}]
    if floor:
        meshes.append({"material": {"type": "phong", "kd": [0.3, 0.3, 0.3], "ks": [0.4, 0.4, 0.4], "exponent": 30.0}, "indices": [0, 2, 1, 0, 3, 2],
                       "P": [-3, -1.2, -3, 3, -1.2, -3, 3, -1.2, 3, -3, -1.2, 3]})
    if area_light:
        meshes.append({"material": {"type": "diffuse", "kd": [0, 0, 0]}, "emission": [5, 5, 5], "indices": [0, 1, 2, 0, 2, 3],
                       "P": [-0.5, 2.5, -0.5, 0.5, 2.5, -0.5, 0.5, 2.5, 0.5, -0.5, 2.5, 0.5]})
    txt = json.dumps({"camera": {"width": w, "height": h, "fov": 40, "to_world": [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, -1, 0, 0, 0, 5, 1]},
                      "environment": [0.8, 0.9, 1.0], "meshes": meshes})
    sc = SceneLoaderManager().load_string(txt, "json")
    if img is not None:
        sc.set_environment_texture(sc.add_bitmap_texture(img))
    return sc


def _dirs(n, seed):
    d = np.random.default_rng(seed).normal(size=(n, 3))
    return (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)


# ---- (1) the oracle against the model ------------------------------------------------------------------------------------
def test_spec_atan2_and_acos_against_numpy():
    """The SPEC-mode atan2 / acos (Cephes kernels in f32) are within 3 ulp of the exact values, cover all four
    quadrants / the axes, and the device copy (tests/emu) returns the same bits."""
    rng = np.random.default_rng(1)
    xy = np.concatenate([rng.normal(size=(4000, 2)), [[0, 1], [0, -1], [1, 0], [-1, 0], [-1, -0.0], [0, 0], [1e-20, -1], [-3, 1e-12]]]).astype(np.float32)
    for x, y in xy:
        got, exact = ob.spec_atan2(y, x), math.atan2(float(y), float(x))
        assert abs(got - exact) <= 3 * float(np.spacing(np.float32(abs(exact)))) + 1e-37, (x, y, got, exact)
        assert got == eb.lib().emu_spec_atan2(float(y), float(x))
    for c in np.concatenate([np.linspace(-1, 1, 4001), [-1.0, 1.0, 0.5, -0.5, 1e-5, -1e-5]]).astype(np.float32):
        got, exact = ob.spec_acos(c), math.acos(float(c))
        assert abs(got - exact) <= 3 * float(np.spacing(np.float32(exact))) + 4e-7 * (abs(c) > 0.999), (c, got, exact)
        assert got == eb.lib().emu_spec_acos(float(c))


@pytest.mark.parametrize("mode", [ob.MATH_SPEC, ob.MATH_LIBM])
def test_env_pdf_is_normalised_and_eval_reads_the_texel(mode):
    """pdf(d) = P(texel) / (2 pi^2 sin theta) integrates to 1 over the sphere; eval(d) is the texel that d's (phi, theta) falls in
    (nearest texel, Bitmap::pixel_uv); black texels have pdf 0."""
    img = sky_image()
    H, W, _ = img.shape
    osc = ob.OracleScene(env_scene(img))
    tot = 0.0
    for j in range(H):
        for i in range(W):
            th, ph = (j + 0.5) * math.pi / H, (i + 0.5) * 2 * math.pi / W
            d = np.float32([math.sin(th) * math.cos(ph), math.sin(th) * math.sin(ph), math.cos(th)])
            rgb, pdf = osc.env_eval_pdf(d, mode)
            assert np.array_equal(rgb, img[j, i])
            assert (pdf == 0.0) == (img[j, i].sum() == 0.0)
            # solid angle of the texel with pdf constant over it up to the 1 / sin(theta) factor: integrate sin analytically
            tot += pdf * math.sin(th) * (2 * math.pi / W) * (math.pi / H)
    assert tot == pytest.approx(1.0, rel=2e-3)


def test_env_samples_follow_the_pdf():
    """sample_direction(u) returns (d, colour, pdf) with colour = eval(d), pdf = pdf(d) (same texel), and E[1 / pdf] over the samples
    = the solid angle of the support (4 pi minus the black texels)."""
    img = sky_image(seed=2)
    H, W, _ = img.shape
    osc = ob.OracleScene(env_scene(img))
    rng = np.random.default_rng(3)
    inv = []
    for u0, u1 in rng.random((4000, 2)).astype(np.float32):
        d, rgb, pdf = osc.env_sample(u0, u1)
        assert abs(np.linalg.norm(d) - 1) < 1e-5 and pdf > 0
        inv.append(1.0 / pdf)
        # Quirk kept from the reference (emitter.rs:364-365): the continuous sample is clamped to [0, size - 1], so every sample of the
        # LAST column / row collapses onto that texel's first edge -- a direction whose own lookup may land in the neighbour texel.
        col, row = (math.atan2(d[1], d[0]) % (2 * math.pi)) / (2 * math.pi) * W, math.acos(max(-1.0, min(1.0, float(d[2])))) / math.pi * H
        if col > W - 1 - 1e-3 or row > H - 1 - 1e-3 or abs(col - round(col)) < 1e-3 or abs(row - round(row)) < 1e-3 or row < 0.05:  # (f32 acos next to the pole)
            continue
        rgb2, pdf2 = osc.env_eval_pdf(d)
        assert np.array_equal(rgb, rgb2) and pdf2 == pytest.approx(pdf, rel=2e-4)
    th = (np.arange(H) + 0.5) * math.pi / H
    # (the last row counts with the sine of its first edge: the clamp quirk above)
    support = sum(2 * math.pi / W * ((math.cos(th[j] - 0.5 * math.pi / H) - math.cos(th[j] + 0.5 * math.pi / H)) if j < H - 1 else math.pi / H * math.sin((H - 1) * math.pi / H))
                  for j in range(H) for i in range(W) if img[j, i].sum() > 0)
    assert np.mean(inv) == pytest.approx(support, rel=0.05)


def test_constant_image_equals_the_constant_environment():
    """An image whose texels all hold c is the constant environment c: with BSDF sampling only the two renders are identical bit for bit
    (eval is c either way; light sampling draws its numbers but contributes nothing), with MIS they agree statistically."""
    c = np.float32([0.8, 0.9, 1.0])
    const, tex = env_scene(None, 24, 24), env_scene(np.tile(c, (4, 8, 1)).astype(np.float32), 24, 24)
    a, sa = ob.OracleScene(const).render(_abi.path_desc(strategy=_abi.RL_STRATEGY_BSDF), 8, seed=3, cfg=ob.config(**STREAM))
    b, sb = ob.OracleScene(tex).render(_abi.path_desc(strategy=_abi.RL_STRATEGY_BSDF), 8, seed=3, cfg=ob.config(**STREAM))
    assert sa.segments == sb.segments and np.array_equal(a, b)
    a, _ = ob.OracleScene(const).render(_abi.path_desc(max_depth=4), 64, seed=3, cfg=ob.config(**STREAM))
    b, _ = ob.OracleScene(tex).render(_abi.path_desc(max_depth=4), 64, seed=3, cfg=ob.config(**STREAM))
    assert abs(a.mean() - b.mean()) < 0.02 * a.mean()


def test_strategies_agree_under_an_environment_texture():
    """BSDF-only, emitter-only and MIS estimators converge to the same image (a bright sun texel: light sampling matters)."""
    osc = ob.OracleScene(env_scene(sky_image(), 16, 16, floor=True))
    imgs = [osc.render(_abi.path_desc(strategy=s, max_depth=3), 600, seed=1, cfg=ob.config(**STREAM))[0] for s in (_abi.RL_STRATEGY_ALL, _abi.RL_STRATEGY_BSDF, _abi.RL_STRATEGY_EMITTER)]
    for other in imgs[1:]:
        assert abs(other.mean() - imgs[0].mean()) < 0.06 * imgs[0].mean()


# ---- (2) device arithmetic == oracle, bit for bit ---------------------------------------------------------------------------
def test_device_env_functions_bit_exact():
    sc = env_scene(sky_image(seed=5))
    esc, osc = eb.EmuScene(sc), ob.OracleScene(sc)
    for d in np.concatenate([_dirs(3000, 7), np.float32([[0, 0, 1], [0, 0, -1], [1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [-1, -1e-9, 0]])]):
        (ce, pe), (co, po) = esc.env_eval_pdf(d), osc.env_eval_pdf(d)
        assert np.array_equal(ce, co) and pe == po, d
    for u0, u1 in np.concatenate([np.random.default_rng(8).random((3000, 2)), [[0, 0], [0.999999, 0.999999], [0, 0.5], [0.5, 0]]]).astype(np.float32):
        (de, ce, pe), (do, co, po) = esc.env_sample(u0, u1), osc.env_sample(u0, u1)
        assert np.array_equal(de, do) and np.array_equal(ce, co) and pe == po, (u0, u1)


@pytest.mark.parametrize("integ", [_abi.path_desc(), _abi.path_desc(strategy=_abi.RL_STRATEGY_BSDF), _abi.path_desc(strategy=_abi.RL_STRATEGY_EMITTER),
                                   _abi.path_desc(max_depth=4, rr_depth=2), _abi.direct_desc(1, 1), _abi.direct_desc(2, 0), _abi.direct_desc(0, 2)],
                         ids=["path", "path-bsdf", "path-emitter", "path-d4", "direct11", "direct20", "direct02"])
def test_environment_texture_render_bit_exact(integ):
    sc = env_scene(sky_image(), 32, 32, area_light=True, floor=True)
    ie, se = eb.EmuScene(sc).render(integ, 6, seed=4)
    io, so = ob.OracleScene(sc).render(integ, 6, seed=4, cfg=ob.config(**STREAM))
    assert np.isfinite(io).all() and io.mean() > 0.01
    assert (se.segments, se.hits, se.shadow_rays) == (so.segments, so.hits, so.shadow_rays)
    assert np.array_equal(ie, io)


def test_environment_texture_stream_estimator_equals_graph_and_math_modes_agree():
    sc = env_scene(sky_image(), 24, 24, floor=True)
    osc = ob.OracleScene(sc)
    integ = _abi.path_desc(max_depth=5)
    a, sa = osc.render(integ, 16, seed=2, cfg=ob.config(**STREAM))
    b, sb = osc.render(integ, 16, seed=2, cfg=ob.config(estimator=ob.EST_GRAPH, accel_mode=ob.ACCEL_BVH))
    assert (sa.segments, sa.shadow_rays) == (sb.segments, sb.shadow_rays) and rel_l2(a, b) < 1e-6
    c, _ = osc.render(integ, 16, seed=2, cfg=ob.config(math_mode=ob.MATH_LIBM, **STREAM))
    assert rel_l2(c, a) < 0.02  # texel lookups can flip at texel borders between libm and spec atan2 / acos: rare, bounded


def test_cornell_box_under_a_sky():
    """The environment texture seen through the open front of the Cornell box (most light paths never escape)."""
    sc = load_cbox(32, 32)
    sc.set_environment_texture(sc.add_bitmap_texture(sky_image()))
    integ = _abi.path_desc(max_depth=5)
    ie, se = eb.EmuScene(sc).render(integ, 4, seed=1)
    io, so = ob.OracleScene(sc).render(integ, 4, seed=1, cfg=ob.config(**STREAM))
    assert se.segments == so.segments and np.array_equal(ie, io)


# ---- (3) loaders ---------------------------------------------------------------------------------------------------------------
def test_pbrt_infinite_light_with_a_mapname(tmp_path):
    from rustlight_b200.host import save_pfm
    img = sky_image()
    save_pfm(str(tmp_path / "sky.pfm"), img)
    pbrt = open(os.path.join(os.path.dirname(__file__), "..", "data", "cbox.pbrt")).read().replace("WorldBegin", 'WorldBegin\nLightSource "infinite" "string mapname" "sky.pfm"', 1)
    (tmp_path / "c.pbrt").write_text(pbrt)
    sc = SceneLoaderManager().load(str(tmp_path / "c.pbrt"))
    sc.set_resolution(24, 24)
    ref = load_cbox(24, 24)
    ref.set_environment_texture(ref.add_bitmap_texture(np.abs(img)))
    integ = _abi.path_desc(max_depth=3)
    a, _ = ob.OracleScene(sc).render(integ, 2, seed=1, cfg=ob.config(**STREAM))
    b, _ = ob.OracleScene(ref).render(integ, 2, seed=1, cfg=ob.config(**STREAM))
    assert np.array_equal(a, b)
    back = SceneLoaderManager().load_string(sc.to_json(), "json")  # JSON round trip keeps the environment texture
    c, _ = ob.OracleScene(back).render(integ, 2, seed=1, cfg=ob.config(**STREAM))
    assert np.array_equal(a, c)
    (tmp_path / "bad.pbrt").write_text(pbrt.replace('"string mapname" "sky.pfm"', '"string mapname" "sky.pfm" "rgb scale" [2 2 2]'))
    with pytest.raises(SceneError):
        SceneLoaderManager().load(str(tmp_path / "bad.pbrt"))  # the reference asserts scale == 1 with a mapname (scene_loader.rs:260-262)


def test_environment_texture_must_be_a_bitmap():
    sc = env_scene(None)
    with pytest.raises(SceneError):
        sc.set_environment_texture(1)
