"""BSDFMetal / BSDFGlass / BSDFSubstrate + MicrofacetDistribution (bsdfs/{metal,glass,substrate,distribution,utils}.rs).

Three layers, like the rest of the suite: (1) the oracle against closed forms and identities of the model,
(2) the device arithmetic (tests/emu) against the oracle, bit for bit, (3) whole renders of a Cornell box that
mixes every material kind: emulator == oracle (stream estimator) bit for bit, stream ~= graph estimator."""
import math

import numpy as np
import pytest

import emu_binding as eb
from conftest import load_cbox, mixed_cbox, rel_l2
from oracle import binding as ob
from rustlight_b200 import _abi
from rustlight_b200.host import (material_glass, material_metal, material_mirror, material_phong, material_substrate)

STREAM = dict(estimator=ob.EST_STREAM, accel_mode=ob.ACCEL_BVH)

MIRROR = material_mirror((0.9, 0.8, 0.7))
GOLD = material_metal((1, 1, 1), (0.143, 0.375, 1.442), (3.983, 2.386, 1.603), "ggx", 0.12)
ALU_B = material_metal((0.9, 0.9, 0.9), (1.657, 0.880, 0.521), (9.224, 6.270, 4.837), "beckmann", 0.25)
GLASS = material_glass((1, 1, 1), (0.95, 0.97, 1.0), 1.5046, 1.000277)
COAT = material_substrate((0.4, 0.25, 0.1), (0.05, 0.05, 0.05), "ggx", 0.1)
COAT_B = material_substrate((0.2, 0.3, 0.4), (0.08, 0.08, 0.08), "beckmann", 0.3)
COAT_DELTA = material_substrate((0.5, 0.5, 0.5), (0.04, 0.04, 0.04), None, 0.0)
ALL = [MIRROR, GOLD, ALU_B, GLASS, COAT, COAT_B, COAT_DELTA]
NAMES = ["mirror", "gold_ggx", "alu_beckmann", "glass", "substrate_ggx", "substrate_beckmann", "substrate_delta"]


def _dir(theta, phi):
    return np.float32([math.sin(theta) * math.cos(phi), math.sin(theta) * math.sin(phi), math.cos(theta)])


# ---- (1) oracle vs the model -------------------------------------------------------------------------
def test_flags():
    want = {"mirror": (True, True), "gold_ggx": (True, False), "alu_beckmann": (True, False), "glass": (False, True),
            "substrate_ggx": (True, False), "substrate_beckmann": (True, False), "substrate_delta": (True, True)}
    for m, n in zip(ALL, NAMES):
        f = ob.bsdf_flags(m)
        assert (f["twosided"], f["smooth"]) == want[n], n
        assert eb.bsdf_flags(m) == f, n


def test_mirror_is_a_conductor_with_eta_1_k_0():
    """pbrt "mirror" -> BSDFMetal{eta: 1, k: 0} (bsdfs/mod.rs:349-357): fresnel_conductor degenerates to
    Rs = (1-c)/(1+c), Rp = Rs ((1-c)/(1+c))... evaluated here in float64 from utils.rs:78-100; the lobe is the mirror direction."""
    for c in (0.1, 0.5, 0.9, 1.0):
        wi = _dir(math.acos(c), 0.7)
        ok, w, d, pdf, disc = ob.bsdf_sample_ex(MIRROR, wi, 0.3, 0.6)
        assert ok and disc and pdf == 1.0 and np.allclose(d, [-wi[0], -wi[1], wi[2]])
        c = float(wi[2])
        c2, s2 = c * c, 1 - c * c
        a2pb2 = math.sqrt(c2 * c2)
        a = math.sqrt(0.5 * (a2pb2 + c2))
        t1, t2 = a2pb2 + c2, a * 2 * c2
        rs = (t1 - t2) / (t1 + t2)
        t3, t4 = a2pb2 * c2 + s2 * s2, t2 * s2
        rp = rs * (t3 - t4) / (t3 + t4)
        assert np.allclose(w, np.float32([0.9, 0.8, 0.7]) * 0.5 * (rp + rs), rtol=2e-5, atol=1e-7)


def test_glass_fresnel_and_snell():
    eta = GLASS.ior
    n_refl = 0
    rng = np.random.default_rng(1)
    for ct in (0.05, 0.3, 0.7, 1.0, -0.05, -0.5, -0.9):
        wi = _dir(math.acos(ct), 1.1)
        # Fresnel (unpolarised) in float64
        scale = 1 / eta if ct > 0 else eta
        ct2 = 1 - (1 - wi[2] ** 2) * scale * scale
        if ct2 <= 0:
            F = 1.0
        else:
            ci, cT = abs(float(wi[2])), math.sqrt(ct2)
            rs = (ci - eta * cT) / (ci + eta * cT)
            rp = (eta * ci - cT) / (eta * ci + cT)
            F = 0.5 * (rs * rs + rp * rp)
        for s0 in rng.random(20):
            ok, w, d, pdf, disc = ob.bsdf_sample_ex(GLASS, wi, float(s0), 0.5)
            assert ok and disc and pdf == pytest.approx(F, rel=2e-5, abs=1e-7)
            if s0 <= pdf:
                n_refl += 1
                assert np.allclose(d, [-wi[0], -wi[1], wi[2]]) and np.array_equal(w, np.float32([1, 1, 1]))
            else:
                assert np.array_equal(w, np.float32(list(GLASS.kt)))  # Importance transport: factor 1 (glass.rs:95-105)
                assert d[2] * wi[2] < 0 and abs(np.linalg.norm(d) - 1) < 1e-5
                # Snell: sin_t = sin_i * (1/eta entering, eta leaving)
                assert math.hypot(d[0], d[1]) == pytest.approx(math.hypot(wi[0], wi[1]) * scale, rel=1e-5)
    assert n_refl > 5
    # total internal reflection from inside at a grazing angle
    ok, w, d, pdf, disc = ob.bsdf_sample_ex(GLASS, _dir(math.radians(100), 0.0), 0.999, 0.5)
    assert ok and pdf == 1.0 and d[2] < 0


@pytest.mark.parametrize("mat", [GOLD, ALU_B])
def test_microfacet_normal_distribution(mat):
    """D(m) cos(theta_m) integrates to 1; sample() returns exactly pdf(m) = D(m) cos(theta_m) of the normal it
    draws -- seen through BSDFMetal: the sampled `pdf` is the pdf of the NORMAL (metal.rs:64, a quirk that is kept),
    while BSDFMetal::pdf() is D cos / (4 |wo.h|)."""
    wi = _dir(0.4, 0.3)
    for s0, s1 in np.random.default_rng(2).random((200, 2)):
        ok, w, d, pdf, disc = ob.bsdf_sample_ex(mat, wi, float(s0), float(s1))
        if not ok:
            continue
        assert not disc and d[2] > 0
        h = (wi + d) / np.linalg.norm(wi + d)
        assert np.allclose(d, -wi + 2 * np.dot(wi, h) * h, atol=2e-6)  # reflect_vector
        p_wo = ob.bsdf_pdf(mat, wi, d)
        assert p_wo == pytest.approx(pdf / (4 * abs(float(np.dot(d, h)))), rel=2e-4)
    # hemispherical integral of pdf(wo) for normal incidence equals 1 minus the mass reflected below the horizon (none here)
    nt, nph = 600, 64
    tot = 0.0
    wi = np.float32([0, 0, 1])
    for t in (np.arange(nt) + 0.5) * (math.pi / 2) / nt:
        tot += sum(ob.bsdf_pdf(mat, wi, _dir(t, p)) for p in (np.arange(nph) + 0.5) * 2 * math.pi / nph) * math.sin(t)
    tot *= (math.pi / 2 / nt) * (2 * math.pi / nph)
    assert tot == pytest.approx(1.0, abs=2e-2)


@pytest.mark.parametrize("mat", [COAT, COAT_B, COAT_DELTA])
def test_substrate_weight_is_eval_over_pdf_and_lobes(mat):
    wi = _dir(0.6, -0.4)
    n_spec = 0
    for s0, s1 in np.random.default_rng(3).random((300, 2)):
        ok, w, d, pdf, disc = ob.bsdf_sample_ex(mat, wi, float(s0), float(s1))
        if not ok:
            continue
        assert d[2] > 0 and pdf > 0
        if disc:  # only the distribution-less substrate has a delta lobe: pdf 1/2, weight = schlick / (1/2)
            n_spec += 1
            assert mat.microfacet == _abi.RL_MICROFACET_NONE and s0 >= 0.5 and pdf == 0.5
            rs = np.float32(list(mat.ks))
            assert np.allclose(w, (rs + (1 - rs) * (1 - wi[2]) ** 5) / 0.5, rtol=1e-5)
        else:
            assert pdf == pytest.approx(ob.bsdf_pdf(mat, wi, d), rel=1e-6)
            assert np.allclose(w, ob.bsdf_eval(mat, wi, d) / np.float32(pdf), rtol=1e-6)
    assert (n_spec > 50) == (mat.microfacet == _abi.RL_MICROFACET_NONE)


def test_math_modes_agree_on_the_new_lobes():
    wi = _dir(0.5, 0.2)
    for mat in (GOLD, ALU_B, COAT, COAT_B):
        for s0, s1 in np.random.default_rng(5).random((50, 2)):
            a = ob.bsdf_sample_ex(mat, wi, float(s0), float(s1), ob.MATH_LIBM)
            b = ob.bsdf_sample_ex(mat, wi, float(s0), float(s1), ob.MATH_SPEC)
            assert a[0] == b[0]
            if a[0]:
                assert np.allclose(a[1], b[1], rtol=2e-5) and np.allclose(a[2], b[2], atol=2e-6) and a[3] == pytest.approx(b[3], rel=2e-5)


# ---- (2) device arithmetic == oracle, bit for bit ----------------------------------------------------------
@pytest.mark.parametrize("mat,name", list(zip(ALL, NAMES)))
def test_device_bsdf_bit_exact(mat, name):
    rng = np.random.default_rng(11)
    n_ok = 0
    for i in range(600):
        wi = _dir(math.acos(rng.uniform(-1 if name == "glass" else -0.2, 1)), rng.uniform(0, 2 * math.pi))
        if i % 50 == 0:
            wi = np.float32([0, 0, 1])
        s0, s1 = float(np.float32(rng.random())), float(np.float32(rng.random()))
        a = ob.bsdf_sample_ex(mat, wi, s0, s1)
        b = eb.bsdf_sample_ex(mat, wi, s0, s1)
        assert a[0] == b[0], (name, wi, s0, s1)
        if not a[0]:
            continue
        n_ok += 1
        assert a[4] == b[4] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]) and a[3] == b[3], (name, wi, s0, s1)
        if not ob.bsdf_flags(mat)["smooth"]:
            wo = _dir(math.acos(rng.uniform(-0.1, 1)), rng.uniform(0, 2 * math.pi))
            for o in (a[2], wo):
                pa, pb = ob.bsdf_pdf(mat, wi, o), eb.bsdf_pdf(mat, wi, o)
                assert pa == pb or (math.isnan(pa) and math.isnan(pb)), (name, wi, o)
                assert np.array_equal(ob.bsdf_eval(mat, wi, o), eb.bsdf_eval(mat, wi, o), equal_nan=True), (name, wi, o)
    assert n_ok > 300


# ---- (3) renders ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kw", [dict(), dict(strategy=_abi.RL_STRATEGY_BSDF), dict(strategy=_abi.RL_STRATEGY_EMITTER), dict(max_depth=5, rr_depth=3)])
def test_mixed_scene_path_bit_exact(kw):
    sc = mixed_cbox()
    integ = _abi.path_desc(**kw)
    ie, se = eb.EmuScene(sc).render(integ, 6, seed=4)
    io, so = ob.OracleScene(sc).render(integ, 6, seed=4, cfg=ob.config(**STREAM))
    assert np.isfinite(io).all() and io.mean() > 0.01
    assert (se.segments, se.hits, se.shadow_rays) == (so.segments, so.hits, so.shadow_rays)
    assert np.array_equal(ie, io)


@pytest.mark.parametrize("nb,nl", [(1, 1), (2, 0), (0, 2)])
def test_mixed_scene_direct_bit_exact(nb, nl):
    sc = mixed_cbox()
    integ = _abi.direct_desc(nb, nl)
    ie, se = eb.EmuScene(sc).render(integ, 4, seed=9)
    io, so = ob.OracleScene(sc).render(integ, 4, seed=9, cfg=ob.config(**STREAM))
    assert (se.segments, se.hits, se.shadow_rays) == (so.segments, so.hits, so.shadow_rays)
    assert np.array_equal(ie, io)


def test_mixed_scene_stream_estimator_equals_graph():
    """path.rs builds the vertex/edge graph and recurses; the streaming form must take the same decisions (same ray
    counts) and differ only by re-association of products -- now including PDF::Discrete edges and smooth vertices."""
    sc = mixed_cbox(40, 40)
    osc = ob.OracleScene(sc)
    for kw in (dict(), dict(strategy=_abi.RL_STRATEGY_BSDF), dict(strategy=_abi.RL_STRATEGY_EMITTER)):
        integ = _abi.path_desc(**kw)
        a, sa = osc.render(integ, 8, seed=2, cfg=ob.config(estimator=ob.EST_GRAPH, accel_mode=ob.ACCEL_BVH))
        b, sb = osc.render(integ, 8, seed=2, cfg=ob.config(**STREAM))
        assert (sa.segments, sa.shadow_rays) == (sb.segments, sb.shadow_rays)
        assert rel_l2(b, a) < 1e-6


def test_strategies_agree_statistically_on_glossy_scene():
    """BSDF-only, emitter-only and MIS estimators of a scene WITHOUT delta lobes converge to the same image."""
    sc = load_cbox(24, 24)
    for i, m in enumerate([COAT, COAT_B, GOLD, ALU_B, COAT]):
        sc.set_material(i, m)
    osc = ob.OracleScene(sc)
    imgs = [osc.render(_abi.path_desc(strategy=s, max_depth=4), 600, seed=1, cfg=ob.config(**STREAM))[0]
            for s in (_abi.RL_STRATEGY_ALL, _abi.RL_STRATEGY_BSDF, _abi.RL_STRATEGY_EMITTER)]
    m = [float(i.mean()) for i in imgs]
    assert m[1] == pytest.approx(m[0], rel=0.06) and m[2] == pytest.approx(m[0], rel=0.06)


# ---- non-mesh emitters (PointEmitter, DirectionalLight: emitter.rs:96-250) ------------------------------------
def lit_cbox(w=48, h=48, keep_area_light=True):
    sc = load_cbox(w, h)
    sc.add_point_light((0.6, 0.5, 0.4), (0.3, 1.2, 0.4))
    sc.add_directional_light((0.8, 0.8, 1.0), (0.3, -1.0, -0.2))
    return sc


def test_point_and_directional_light_sampling():
    """sample_light over [area light, point, directional]: selection by flux (scene.rs:103-111), PDF::Discrete(1) * pdf_sel,
    point weight I / d^2 / pdf_sel, directional p = x - 1.1 R dir with R from Scene.bsphere (meshes + camera)."""
    sc = lit_cbox()
    osc = ob.OracleScene(sc)
    x = np.float32([0.1, 0.5, 0.2])
    seen = set()
    for r_sel in np.linspace(0.001, 0.999, 40):
        rec = osc.sample_light(x, float(r_sel), 0.3, 0.4, 0.6)
        e, p, n, d, w, pdf = rec["mesh"], rec["p"], rec["n"], rec["d"], rec["weight"], rec["pdf"]
        seen.add(e)
        if e == -3 and not n.any():  # point light
            dist2 = float(((np.float32([0.3, 1.2, 0.4]) - x) ** 2).sum())
            assert np.allclose(p, [0.3, 1.2, 0.4]) and np.allclose(w * pdf, np.float32([0.6, 0.5, 0.4]) / dist2, rtol=1e-5)
        elif e == -3:  # directional
            dn = np.float32([0.3, -1.0, -0.2]) / np.linalg.norm([0.3, -1.0, -0.2])
            assert np.allclose(n, dn, atol=1e-6) and np.allclose(d, -dn, atol=1e-6) and np.allclose(w * pdf, [0.8, 0.8, 1.0], rtol=1e-5)
            # bounding sphere of the box [-1,1]x[0,2]x[-1,1] united with the camera at z = 6.8: centre (0,1,2.9), R = |(1,1,3.9)|
            R = 1.1 * math.sqrt(1 + 1 + 3.9 ** 2)
            assert np.linalg.norm(p - x) == pytest.approx(R, rel=1e-4)
    assert -3 in seen and any(e >= 0 for e in seen)


@pytest.mark.parametrize("integ", [_abi.path_desc(), _abi.path_desc(strategy=_abi.RL_STRATEGY_EMITTER, max_depth=3), _abi.direct_desc(1, 2)])
def test_delta_lights_render_bit_exact(integ):
    sc = lit_cbox()
    ie, se = eb.EmuScene(sc).render(integ, 6, seed=7)
    io, so = ob.OracleScene(sc).render(integ, 6, seed=7, cfg=ob.config(**STREAM))
    assert (se.segments, se.hits, se.shadow_rays) == (so.segments, so.hits, so.shadow_rays)
    assert np.array_equal(ie, io) and io.mean() > 0.05
    if integ.kind == _abi.RL_INTEGRATOR_PATH:
        ig, sg = ob.OracleScene(sc).render(integ, 6, seed=7, cfg=ob.config(estimator=ob.EST_GRAPH, accel_mode=ob.ACCEL_BVH))
        assert rel_l2(io, ig) < 1e-6 and sg.segments == so.segments


def test_point_light_only_scene_matches_inverse_square_law():
    """No emissive mesh: a diffuse floor under one point light, `direct -b 0 -l 1`: radiance = kd/pi * I cos / d^2, a
    closed form per pixel (the light is a delta; only the pixel jitter moves the hit point)."""
    import json
    from rustlight_b200 import SceneLoaderManager
    kd, I, hgt = 0.6, 3.0, 1.5
    txt = json.dumps({"camera": {"width": 16, "height": 16, "fov": 30, "to_world": [1, 0, 0, 0, 0, 0, -1, 0, 0, -1, 0, 0, 0, 4, 0, 1]},
                      "lights": [{"type": "point", "intensity": [I, I, I], "position": [0.0, hgt, 0.0]}],
                      "meshes": [{"material": {"type": "diffuse", "kd": [kd] * 3}, "indices": [0, 1, 2, 0, 2, 3],
                                  "P": [-5, 0, -5, -5, 0, 5, 5, 0, 5, 5, 0, -5]},
                                 # a far-away sliver that lifts the root box above the light: BVHAccel::visible answers "occluded"
                                 # for any segment that leaves the root box within tnear (accel.rs:338-340), which a flat scene
                                 # would hit for the near-vertical shadow rays under the light
                                 {"material": {"type": "diffuse", "kd": [0.5] * 3}, "indices": [0, 1, 2], "P": [40, 5, 40, 41, 5, 40, 40, 5, 41]}]})
    sc = SceneLoaderManager().load_string(txt, "json")
    assert sc.desc.contents.nlights == 1
    osc = ob.OracleScene(sc)
    img, st = osc.render(_abi.direct_desc(0, 1), 64, seed=1, cfg=ob.config(**STREAM))
    pg, tg = osc.primary_hits(ob.ACCEL_BVH)
    assert (pg != 0xFFFFFFFF).all()
    # hit points of the pixel centres: camera at (0,4,0) looking down
    o = np.float32([0, 4, 0])
    ys, xs = np.mgrid[0:16, 0:16]
    rays = np.array([osc.camera_generate(float(x) + 0.5, float(y) + 0.5)[1] for y, x in zip(ys.ravel(), xs.ravel())])
    hit = o + rays * tg.reshape(-1, 3)[:, :1]
    d2 = (hit[:, 0] ** 2 + hgt ** 2 + hit[:, 2] ** 2)
    want = kd / math.pi * I * (hgt / np.sqrt(d2)) / d2
    assert np.allclose(img[..., 0].ravel(), want, rtol=0.03)
    ie, _ = eb.EmuScene(sc).render(_abi.direct_desc(0, 1), 64, seed=1)
    assert np.array_equal(ie, img)


# ---- BSDFColor::{Bitmap, Checkerbord, Grid} on the diffuse slot (bsdfs/mod.rs:31-101) ----------------------------
def _tex_expected(kind, uv, **kw):
    """Independent numpy restatement of BSDFColor::color for an array of uv (float32 arithmetic where it matters)."""
    u, v = uv[:, 0].astype(np.float32), uv[:, 1].astype(np.float32)
    if kind == "bitmap":
        img = kw["img"]
        h, w, _ = img.shape
        x = (np.mod(np.mod(u, 1) + 1, 1) * np.float32(w)).astype(np.int64)
        y = (np.mod(np.mod(v, 1) + 1, 1) * np.float32(h)).astype(np.int64)
        return img[np.clip(y, 0, h - 1), np.clip(x, 0, w - 1)]
    c0, c1 = np.float32(kw["c0"]), np.float32(kw["c1"])
    ox, oy, sx, sy = kw["offset"] + kw["scale"]
    if kind == "checkerboard":
        a, b = u * np.float32(sx) + np.float32(ox), v * np.float32(sy) + np.float32(oy)
        x = 2 * (np.fmod(np.trunc(a * 2), 2)).astype(np.int64) - 1
        y = 2 * (np.fmod(np.trunc(b * 2), 2)).astype(np.int64) - 1
        return np.where((x * y == 1)[:, None], c0, c1)
    a, b = u * np.float32(sx) + np.float32(ox), (v + np.float32(sy)) + np.float32(oy)  # grid: uv.y + scale.y (bsdfs/mod.rs:84)
    x, y = a - np.floor(a), b - np.floor(b)
    x, y = np.where(x > 0.5, x - 1, x), np.where(y > 0.5, y - 1, y)
    return np.where(((np.abs(x) < kw["lw"]) | (np.abs(y) < kw["lw"]))[:, None], c0, c1)


def _textured_floor(kind, **kw):
    """A [-1,1]^2 floor with uv in [0,1]^2 under a point light, seen from above; returns (scene, uv of the pixel centres)."""
    import json
    from rustlight_b200 import SceneLoaderManager
    tex = {"type": kind}
    if kind == "bitmap":
        img = kw["img"]
        tex.update(width=img.shape[1], height=img.shape[0], pixels=[float(x) for x in img.ravel()])
    else:
        tex.update(color0=list(kw["c0"]), color1=list(kw["c1"]), offset=list(kw["offset"]), scale=list(kw["scale"]))
        if kind == "grid":
            tex["line_width"] = kw["lw"]
    txt = json.dumps({"camera": {"width": 48, "height": 48, "fov": 28, "to_world": [1, 0, 0, 0, 0, 0, -1, 0, 0, -1, 0, 0, 0, 4, 0, 1]},
                      "textures": {"t": tex}, "lights": [{"type": "point", "intensity": [3, 3, 3], "position": [0.0, 1.5, 0.0]}],
                      "meshes": [{"material": {"type": "diffuse", "kd_texture": "t"}, "indices": [0, 1, 2, 0, 2, 3],
                                  "P": [-1, 0, -1, 1, 0, -1, 1, 0, 1, -1, 0, 1], "uv": [0, 0, 1, 0, 1, 1, 0, 1]},
                                 {"material": {"type": "diffuse", "kd": [0.5] * 3}, "indices": [0, 1, 2], "P": [40, 5, 40, 41, 5, 40, 40, 5, 41]}]})
    return SceneLoaderManager().load_string(txt, "json")


TEX_CASES = [("checkerboard", dict(c0=(0.9, 0.1, 0.1), c1=(0.1, 0.2, 0.9), offset=(0.13, 0.0), scale=(2.0, 1.5))),
             ("grid", dict(c0=(0.05, 0.05, 0.05), c1=(0.8, 0.7, 0.6), offset=(0.0, 0.1), scale=(3.0, 0.25), lw=0.06)),
             ("bitmap", dict(img=np.random.default_rng(5).random((5, 7, 3)).astype(np.float32)))]


@pytest.mark.parametrize("kind,kw", TEX_CASES)
def test_texture_lookup_semantics(kind, kw):
    """Oracle image of the textured floor / image of the same floor with kd = 1 == the texture value at the pixel's uv,
    for pixels whose footprint stays inside one texel / cell (checked against an independent numpy restatement)."""
    sc = _textured_floor(kind, **kw)
    osc = ob.OracleScene(sc)
    integ = _abi.direct_desc(0, 1)
    img, _ = osc.render(integ, 32, seed=3, cfg=ob.config(**STREAM))
    pg, tg = osc.primary_hits(ob.ACCEL_BVH)
    floor = (pg.ravel() <= 1)
    assert floor.mean() > 0.5
    ys, xs = np.mgrid[0:48, 0:48]
    want = np.zeros((48 * 48, 3), np.float32)
    stable = np.ones(48 * 48, bool)
    probes = [(0.5, 0.5)] + [(a, b) for a in (0.01, 0.25, 0.5, 0.75, 0.99) for b in (0.01, 0.25, 0.5, 0.75, 0.99)]
    for dx, dy in probes:  # centre first, then a 5 x 5 lattice over the pixel footprint
        rays = np.array([osc.camera_generate(float(x) + dx, float(y) + dy)[1] for y, x in zip(ys.ravel(), xs.ravel())])
        t = 4.0 / -rays[:, 1]
        hit = np.float32([0, 4, 0]) + rays * t[:, None]
        uv = np.stack([(hit[:, 0] + 1) / 2, (hit[:, 2] + 1) / 2], axis=1)
        val = _tex_expected(kind, uv, **kw)
        if (dx, dy) == (0.5, 0.5):
            want = val
        else:
            stable &= (val == want).all(axis=1)
    sel = floor & stable
    assert sel.sum() > 400
    # same geometry with a constant kd = 1: divide out the lighting
    sc2 = _textured_floor("checkerboard", c0=(1, 1, 1), c1=(1, 1, 1), offset=(0, 0), scale=(1, 1))
    base, _ = ob.OracleScene(sc2).render(integ, 32, seed=3, cfg=ob.config(**STREAM))
    ratio = img.reshape(-1, 3)[sel] / base.reshape(-1, 3)[sel]
    assert np.allclose(ratio, want[sel], rtol=2e-5, atol=1e-6)


@pytest.mark.parametrize("kind,kw", TEX_CASES)
def test_textured_render_bit_exact(kind, kw):
    sc = _textured_floor(kind, **kw)
    for integ in (_abi.direct_desc(1, 1), _abi.path_desc(max_depth=3)):
        ie, _ = eb.EmuScene(sc).render(integ, 4, seed=2)
        io, _ = ob.OracleScene(sc).render(integ, 4, seed=2, cfg=ob.config(**STREAM))
        assert np.array_equal(ie, io) and io.max() > 0


def test_textured_cornell_box_bit_exact():
    """Textures on the Cornell box (its meshes carry uv): checkerboard floor, bitmap back wall, grid on a substrate box."""
    from rustlight_b200.host import material_diffuse
    sc = load_cbox(48, 48)
    t1 = sc.add_checkerboard_texture((0.8, 0.8, 0.8), (0.1, 0.1, 0.1), (0, 0), (2, 2))
    t2 = sc.add_bitmap_texture(np.random.default_rng(9).random((8, 8, 3)).astype(np.float32))
    t3 = sc.add_grid_texture((0.9, 0.2, 0.2), (0.3, 0.3, 0.3), 0.05, (0, 0), (4, 1))
    sc.set_material(0, material_diffuse(kd_texture=t1))
    sc.set_material(2, material_diffuse(kd_texture=t2))
    m = material_substrate((0, 0, 0), (0.05, 0.05, 0.05), "ggx", 0.2)
    m.kd_texture = t3
    sc.set_material(5, m)
    d = sc.desc.contents
    assert bool(d.meshes[0].UV) and d.ntextures == 3
    integ = _abi.path_desc()
    ie, se = eb.EmuScene(sc).render(integ, 6, seed=4)
    io, so = ob.OracleScene(sc).render(integ, 6, seed=4, cfg=ob.config(**STREAM))
    assert se.segments == so.segments and np.array_equal(ie, io)
    ig, _ = ob.OracleScene(sc).render(integ, 6, seed=4, cfg=ob.config(estimator=ob.EST_GRAPH, accel_mode=ob.ACCEL_BVH))
    assert rel_l2(io, ig) < 1e-6


def test_texture_without_uv_is_black():
    """"Found a texture but no uv coordinate given" -> Color::zero() (bsdfs/mod.rs:36-39): the floor reflects nothing."""
    import json
    from rustlight_b200 import SceneLoaderManager
    sc = _textured_floor("checkerboard", c0=(1, 1, 1), c1=(1, 1, 1), offset=(0, 0), scale=(1, 1))
    j = json.loads(sc.to_json())
    del j["meshes"][0]["uv"]
    sc2 = SceneLoaderManager().load_string(json.dumps(j), "json")
    assert not sc2.desc.contents.meshes[0].UV and sc2.desc.contents.meshes[0].mat.kd_texture == 1
    io, _ = ob.OracleScene(sc2).render(_abi.direct_desc(1, 1), 4, seed=1, cfg=ob.config(**STREAM))
    ie, _ = eb.EmuScene(sc2).render(_abi.direct_desc(1, 1), 4, seed=1)
    assert not io.any() and np.array_equal(ie, io)
    lit, _ = ob.OracleScene(sc).render(_abi.direct_desc(1, 1), 4, seed=1, cfg=ob.config(**STREAM))
    assert lit.mean() > 0.01


# ---- EnvironmentLight with a constant colour (emitter.rs:428-568) ---------------------------------------------------
def _env_scene(w=32, h=32, with_area_light=False):
    """A diffuse sphere-ish blob (an octahedron) floating in a constant environment, seen from outside."""
    import json
    from rustlight_b200 import SceneLoaderManager
    P = [1, 0, 0, -1, 0, 0, 0, 1, 0, 0, -1, 0, 0, 0, 1, 0, 0, -1]
    idx = [0, 2, 4, 2, 1, 4, 1, 3, 4, 3, 0, 4, 2, 0, 5, 1, 2, 5, 3, 1, 5, 0, 3, 5]
    meshes = [{"material": {"type": "diffuse", "kd": [0.6, 0.5, 0.4]}, "indices": idx, "P": P}]
    if with_area_light:
        meshes.append({"material": {"type": "diffuse", "kd": [0, 0, 0]}, "emission": [5, 5, 5], "indices": [0, 1, 2, 0, 2, 3],
                       "P": [-0.5, 2.5, -0.5, 0.5, 2.5, -0.5, 0.5, 2.5, 0.5, -0.5, 2.5, 0.5]})
    txt = json.dumps({"camera": {"width": w, "height": h, "fov": 40, "to_world": [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, -1, 0, 0, 0, 5, 1]},
                      "environment": [0.8, 0.9, 1.0], "meshes": meshes})
    return SceneLoaderManager().load_string(txt, "json")


def test_environment_white_furnace():
    """A convex diffuse body under a uniform environment L: every camera ray sees L on a miss and L * kd / (1 - 0) ... more
    precisely radiance = L * kd after one bounce, because a convex body never sees itself: pixel = L * kd exactly in the
    expectation, for BSDF sampling, light sampling and MIS alike."""
    sc = _env_scene(24, 24)
    osc = ob.OracleScene(sc)
    pg, _ = osc.primary_hits(ob.ACCEL_BVH)
    hit = (pg != 0xFFFFFFFF)
    assert 0.05 < hit.mean() < 0.9
    L, kd = np.float32([0.8, 0.9, 1.0]), np.float32([0.6, 0.5, 0.4])
    for strat in (_abi.RL_STRATEGY_ALL, _abi.RL_STRATEGY_BSDF, _abi.RL_STRATEGY_EMITTER):
        img, _ = osc.render(_abi.path_desc(strategy=strat, max_depth=3), 1500, seed=3, cfg=ob.config(**STREAM))
        inner = np.zeros_like(hit)
        inner[1:-1, 1:-1] = hit[1:-1, 1:-1] & hit[:-2, 1:-1] & hit[2:, 1:-1] & hit[1:-1, :-2] & hit[1:-1, 2:]
        outer = ~hit
        outer[1:-1, 1:-1] &= ~hit[:-2, 1:-1] & ~hit[2:, 1:-1] & ~hit[1:-1, :-2] & ~hit[1:-1, 2:]
        assert np.allclose(img[outer], L, rtol=1e-4)  # the sensor edge is un-weighted in every strategy (path.rs:152-165); f32 sum of 1500 terms
        assert np.allclose(img[inner].mean(axis=0), L * kd, rtol=0.03), strat


@pytest.mark.parametrize("integ", [_abi.path_desc(), _abi.path_desc(strategy=_abi.RL_STRATEGY_BSDF, max_depth=4), _abi.path_desc(strategy=_abi.RL_STRATEGY_EMITTER, max_depth=3),
                                   _abi.direct_desc(1, 1), _abi.direct_desc(2, 0), _abi.direct_desc(0, 2)])
def test_environment_render_bit_exact(integ):
    sc = _env_scene(32, 32, with_area_light=True)
    ie, se = eb.EmuScene(sc).render(integ, 6, seed=5)
    io, so = ob.OracleScene(sc).render(integ, 6, seed=5, cfg=ob.config(**STREAM))
    assert (se.segments, se.hits, se.shadow_rays) == (so.segments, so.hits, so.shadow_rays)
    assert np.array_equal(ie, io) and io.mean() > 0.3
    if integ.kind == _abi.RL_INTEGRATOR_PATH:
        ig, sg = ob.OracleScene(sc).render(integ, 6, seed=5, cfg=ob.config(estimator=ob.EST_GRAPH, accel_mode=ob.ACCEL_BVH))
        assert sg.segments == so.segments and rel_l2(io, ig) < 1e-6


def test_environment_in_the_cornell_box_keeps_the_reference_quirk():
    """BoundingSphere::intersect solves with b = +2 d_p.d (structure.rs:899-903): the distance it returns is the one to the
    sphere BEHIND the ray, so the shadow segment of an environment sample ends after that length and occluders beyond it are not
    seen.  Mirrored verbatim (device == oracle bit for bit); consequence: inside the Cornell box the light-sampling estimator
    leaks environment light through the walls and reads higher than the (unbiased) BSDF-sampling one."""
    sc = load_cbox(24, 24)
    sc.set_environment((0.3, 0.3, 0.4))
    osc = ob.OracleScene(sc)
    integ = _abi.path_desc(max_depth=4)
    ie, se = eb.EmuScene(sc).render(integ, 8, seed=2)
    io, so = osc.render(integ, 8, seed=2, cfg=ob.config(**STREAM))
    assert se.segments == so.segments and np.array_equal(ie, io)
    m = [float(osc.render(_abi.path_desc(strategy=s, max_depth=4), 600, seed=2, cfg=ob.config(**STREAM))[0].mean())
         for s in (_abi.RL_STRATEGY_BSDF, _abi.RL_STRATEGY_EMITTER)]
    assert m[1] > 1.1 * m[0]


def _cbox_textured_on_every_slot(w=48, h=48):
    """Cornell box whose materials carry a texture on every colour slot bsdf_pbrt runs through bsdf_texture_match_pbrt
    (bsdfs/mod.rs:294-386): phong Ks, substrate Ks (+ Kd), mirror Kr, metal eta and k, glass Kr and Kt."""
    from rustlight_b200.host import material_glass, material_metal, material_mirror, material_phong, material_substrate
    sc = load_cbox(w, h)
    rng = np.random.default_rng(21)
    chk = sc.add_checkerboard_texture((0.9, 0.6, 0.3), (0.2, 0.2, 0.25), (0, 0), (3, 3))
    bmp = sc.add_bitmap_texture((0.2 + 0.7 * rng.random((6, 6, 3))).astype(np.float32))
    grid = sc.add_grid_texture((0.95, 0.95, 0.9), (0.4, 0.5, 0.6), 0.08, (0, 0), (3, 2))
    eta_t = sc.add_bitmap_texture((0.1 + 1.4 * rng.random((4, 4, 3))).astype(np.float32))
    k_t = sc.add_bitmap_texture((2.0 + 3.0 * rng.random((4, 4, 3))).astype(np.float32))
    m = material_phong((0.3, 0.3, 0.3), (0.0, 0.0, 0.0), 30.0)
    m.ks_texture = chk
    m.weight_specular = 0.4  # the caller's constant: the reference computes it once from constant colours (bsdfs/mod.rs:518-523)
    sc.set_material(0, m)                                        # floor: phong with a textured Ks
    m = material_substrate((0.0, 0.0, 0.0), (0.0, 0.0, 0.0), "ggx", 0.15)
    m.kd_texture, m.ks_texture = bmp, grid
    sc.set_material(1, m)                                        # ceiling: substrate, both slots textured
    m = material_mirror((0.0, 0.0, 0.0))
    m.ks_texture = grid
    sc.set_material(3, m)                                        # a side wall: mirror with a textured Kr
    m = material_metal((1, 1, 1), (0, 0, 0), (0, 0, 0), "ggx", 0.2)
    m.eta_texture, m.k_texture = eta_t, k_t
    sc.set_material(2, m)                                        # back wall: rough metal, eta and k from bitmaps
    m = material_glass((0, 0, 0), (0, 0, 0), 1.5, 1.0)
    m.ks_texture, m.kt_texture = chk, bmp
    sc.set_material(5, m)                                        # short box: glass with textured Kr and Kt
    return sc


def test_textures_on_every_colour_slot_bit_exact():
    """Device arithmetic (emulator) == oracle with a texture on Ks / Kr / Kt / eta / k, and the stream estimator still equals the
    reference's graph estimator there."""
    sc = _cbox_textured_on_every_slot()
    d = sc.desc.contents
    assert d.ntextures == 5 and d.meshes[2].mat.eta_texture == 4 and d.meshes[5].mat.kt_texture == 2
    stream = {}
    for name, integ in (("path", _abi.path_desc()), ("direct", _abi.direct_desc(1, 1))):
        ie, se = eb.EmuScene(sc).render(integ, 6, seed=9)
        io, so = ob.OracleScene(sc).render(integ, 6, seed=9, cfg=ob.config(**STREAM))
        assert se.segments == so.segments and np.array_equal(ie, io) and io.max() > 0
        stream[name] = io
    ig, _ = ob.OracleScene(sc).render(_abi.path_desc(), 6, seed=9, cfg=ob.config(estimator=ob.EST_GRAPH, accel_mode=ob.ACCEL_BVH))
    assert rel_l2(stream["path"], ig) < 1e-6
    # the textures matter: constant-colour materials give another image
    plain = load_cbox(48, 48)
    ip, _ = ob.OracleScene(plain).render(_abi.path_desc(), 6, seed=9, cfg=ob.config(**STREAM))
    assert rel_l2(ig, ip) > 0.05
    # and they survive the JSON round trip
    from rustlight_b200 import SceneLoaderManager
    sc2 = SceneLoaderManager().load_string(sc.to_json(), "json")
    m2 = sc2.desc.contents.meshes[5].mat
    assert (m2.ks_texture, m2.kt_texture) == (1, 2) and sc2.desc.contents.meshes[2].mat.k_texture == 5
