"""BSDFBlend (bsdfs/blend.rs:3-95): weight * bsdf1 + (1 - weight) * bsdf2 over two rough BSDFs.

Layers as in test_bsdfs.py: (1) the oracle against the identities of the model, (2) the device arithmetic (tests/emu) against
the oracle bit for bit, (3) renders of a Cornell box with blended materials: emulator == oracle (stream estimator), stream ~=
graph estimator, JSON round trip.  The GPU render of the same box is in test_gpu.py."""
import math

import numpy as np
import pytest

import emu_binding as eb
from conftest import _diffuse, blend_triple, blended_cbox, rel_l2
from oracle import binding as ob
from rustlight_b200 import SceneLoaderManager, _abi
from rustlight_b200.host import SceneError, material_glass, material_metal, material_mirror, material_phong, material_substrate

STREAM = dict(estimator=ob.EST_STREAM, accel_mode=ob.ACCEL_BVH)
RED = _diffuse((0.6, 0.1, 0.1))
PHONG = material_phong((0.2, 0.2, 0.2), (0.5, 0.5, 0.5), 30.0)
GOLD = material_metal((1, 1, 1), (0.143, 0.375, 1.442), (3.983, 2.386, 1.603), "ggx", 0.2)
COAT_B = material_substrate((0.2, 0.3, 0.4), (0.08, 0.08, 0.08), "beckmann", 0.3)
PAIRS = [(RED, PHONG, 0.35), (GOLD, RED, 0.5), (COAT_B, PHONG, 0.75), (PHONG, GOLD, 0.0), (RED, COAT_B, 1.0)]
IDS = ["diffuse+phong", "gold+diffuse", "substrate+phong", "weight0", "weight1"]


def _dir(theta, phi):
    return np.float32([math.sin(theta) * math.cos(phi), math.sin(theta) * math.sin(phi), math.cos(theta)])


# ---- (1) the oracle against the model ------------------------------------------------------------------------------
@pytest.mark.parametrize("a,b,w", PAIRS, ids=IDS)
def test_blend_is_the_weighted_sum_of_its_parts(a, b, w):
    """eval = w * eval_1 + (1 - w) * eval_2, pdf likewise (blend.rs:47-76); sample() picks part 1 iff sample.x < w with the
    rescaled number, returns the direction that part samples, the pdf of the whole blend and weight = eval / pdf (blend.rs:10-45)."""
    m = blend_triple(a, b, w)
    assert ob.bsdf_flags(m) == {"twosided": True, "smooth": False}
    rng = np.random.default_rng(3)
    w32 = np.float32(w)
    for _ in range(200):
        wi, wo = _dir(math.acos(rng.uniform(0.05, 1)), rng.uniform(0, 6.28)), _dir(math.acos(rng.uniform(-0.1, 1)), rng.uniform(0, 6.28))
        ea, eb_ = ob.bsdf_eval(a, wi, wo), ob.bsdf_eval(b, wi, wo)
        assert np.array_equal(ob.bsdf_eval(m, wi, wo), ea * w32 + eb_ * (np.float32(1) - w32))
        pa, pb = np.float32(ob.bsdf_pdf(a, wi, wo)), np.float32(ob.bsdf_pdf(b, wi, wo))
        assert np.float32(ob.bsdf_pdf(m, wi, wo)) == pa * w32 + pb * (np.float32(1) - w32)
        s0, s1 = np.float32(rng.random()), np.float32(rng.random())
        first = s0 < w32
        s0p = s0 * (np.float32(1) / w32) if first else (s0 - w32) * (np.float32(1) / (np.float32(1) - w32))
        part = ob.bsdf_sample_ex(a if first else b, wi, float(s0p), float(s1))
        got = ob.bsdf_sample_ex(m, wi, float(s0), float(s1))
        if not part[0]:
            assert not got[0]
            continue
        p = ob.bsdf_pdf(m, wi, part[2])
        assert got[0] == (p != 0.0)
        if got[0]:
            assert np.array_equal(got[2], part[2]) and got[3] == p and not got[4]
            assert np.array_equal(got[1], ob.bsdf_eval(m, wi, part[2]) / np.float32(p))


def test_blend_of_smooth_or_nested_parts_is_rejected():
    """blend.rs:17 asserts !is_smooth() on both parts; the host layer refuses them (and nested blends, textures) when the material is set."""
    from conftest import load_cbox
    sc = load_cbox(16, 16)
    for bad in (material_mirror((0.9, 0.9, 0.9)), material_glass(), material_substrate((0.5, 0.5, 0.5), (0.04, 0.04, 0.04), None, 0.0)):
        with pytest.raises(SceneError):
            sc.set_material_blend(0, RED, bad, 0.5)
    with pytest.raises(SceneError):
        sc.set_material_blend(0, RED, PHONG, 1.5)
    tex = _diffuse((0.5, 0.5, 0.5))
    tex.kd_texture = 1
    with pytest.raises(SceneError):
        sc.set_material_blend(0, RED, tex, 0.5)
    with pytest.raises(SceneError):
        sc.set_material(0, blend_triple(RED, PHONG, 0.5))  # a blend needs its parts: set_material_blend


# ---- (2) device arithmetic == oracle, bit for bit --------------------------------------------------------------------
@pytest.mark.parametrize("a,b,w", PAIRS, ids=IDS)
def test_device_blend_bit_exact(a, b, w):
    m = blend_triple(a, b, w)
    assert eb.bsdf_flags(m) == ob.bsdf_flags(m)
    rng = np.random.default_rng(17)
    n_ok = 0
    for i in range(500):
        wi = _dir(math.acos(rng.uniform(-0.2, 1)), rng.uniform(0, 2 * math.pi))
        if i % 50 == 0:
            wi = np.float32([0, 0, 1])
        s0, s1 = float(np.float32(rng.random())), float(np.float32(rng.random()))
        x, y = ob.bsdf_sample_ex(m, wi, s0, s1), eb.bsdf_sample_ex(m, wi, s0, s1)
        assert x[0] == y[0], (wi, s0, s1)
        wo = _dir(math.acos(rng.uniform(-0.1, 1)), rng.uniform(0, 2 * math.pi))
        for o in ((x[2], wo) if x[0] else (wo,)):
            pa, pb = ob.bsdf_pdf(m, wi, o), eb.bsdf_pdf(m, wi, o)
            assert pa == pb or (math.isnan(pa) and math.isnan(pb)), (wi, o)
            assert np.array_equal(ob.bsdf_eval(m, wi, o), eb.bsdf_eval(m, wi, o), equal_nan=True), (wi, o)
        if x[0]:
            n_ok += 1
            assert x[4] == y[4] and np.array_equal(x[1], y[1]) and np.array_equal(x[2], y[2]) and x[3] == y[3], (wi, s0, s1)
    assert n_ok > 250


# ---- (3) renders -------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("integ", [_abi.path_desc(), _abi.path_desc(strategy=_abi.RL_STRATEGY_BSDF), _abi.path_desc(max_depth=5, rr_depth=3),
                                   _abi.direct_desc(1, 1), _abi.direct_desc(0, 2)], ids=["path", "path-bsdf", "path-d5", "direct11", "direct02"])
def test_blended_scene_bit_exact(integ):
    sc = blended_cbox()
    ie, se = eb.EmuScene(sc).render(integ, 6, seed=4)
    io, so = ob.OracleScene(sc).render(integ, 6, seed=4, cfg=ob.config(**STREAM))
    assert np.isfinite(io).all() and io.mean() > 0.01
    assert (se.segments, se.hits, se.shadow_rays) == (so.segments, so.hits, so.shadow_rays)
    assert np.array_equal(ie, io)


def test_blended_scene_stream_estimator_equals_graph():
    sc = blended_cbox(32, 32)
    integ = _abi.path_desc(max_depth=6)
    a, sa = ob.OracleScene(sc).render(integ, 8, seed=2, cfg=ob.config(**STREAM))
    b, sb = ob.OracleScene(sc).render(integ, 8, seed=2, cfg=ob.config(estimator=ob.EST_GRAPH, accel_mode=ob.ACCEL_BVH))
    assert (sa.segments, sa.hits, sa.shadow_rays) == (sb.segments, sb.hits, sb.shadow_rays)
    assert rel_l2(a, b) < 1e-6


def test_blend_weight_0_and_1_render_like_the_single_part():
    """weight = 1 never samples part 2 and weights it by 0: the image equals the image of part 1 alone up to the rounding of
    eval / pdf instead of the part's own weight (statistically identical, same paths)."""
    from conftest import load_cbox
    integ = _abi.path_desc(max_depth=4)
    one = load_cbox(24, 24)
    one.set_material(0, PHONG)
    both = load_cbox(24, 24)
    both.set_material_blend(0, PHONG, RED, 1.0)
    a, sa = ob.OracleScene(one).render(integ, 8, seed=1, cfg=ob.config(**STREAM))
    b, sb = ob.OracleScene(both).render(integ, 8, seed=1, cfg=ob.config(**STREAM))
    assert sa.segments == sb.segments and rel_l2(a, b) < 1e-5


def test_blend_json_round_trip(tmp_path):
    sc = blended_cbox(16, 16)
    p = tmp_path / "blend.json"
    p.write_text(sc.to_json())
    back = SceneLoaderManager().load(str(p))
    integ = _abi.path_desc(max_depth=3)
    a, _ = ob.OracleScene(sc).render(integ, 2, seed=1, cfg=ob.config(**STREAM))
    b, _ = ob.OracleScene(back).render(integ, 2, seed=1, cfg=ob.config(**STREAM))
    assert np.array_equal(a, b)
    assert back.to_json() == sc.to_json()
