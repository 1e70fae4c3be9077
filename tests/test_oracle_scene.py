"""Scene-level pins of the oracle: camera, emitters, accelerators, estimators, closed forms."""
import json
import math

import numpy as np
import pytest

from conftest import load_cbox, rel_l2, soup_scene
from oracle import binding as ob
from rustlight_b200 import SceneLoaderManager, _abi
from rustlight_b200.host import camera_create


# ---- Camera::new / generate, camera.rs:31-91 -----------------------------------------------------------
def test_camera_host_matches_oracle(cbox):
    cam = cbox.desc.contents.camera
    tw = np.array(cam.to_world, np.float32)
    s2c_host, _ = camera_create(512, 512, 19.5, tw)
    s2c_orc = ob.camera_new(512, 512, 19.5, tw)
    assert np.array_equal(np.array(cam.sample_to_camera, np.float32), s2c_host)
    assert np.allclose(s2c_host, s2c_orc, rtol=2e-6, atol=1e-9)  # host-only matrix inverse: last-bit slack


def test_camera_generate(cbox_oracle):
    o, d = cbox_oracle.camera_generate(256.0, 256.0)
    assert np.allclose(o, [0, 1, 6.8], atol=1e-6) and np.allclose(d, [0, 0, -1], atol=1e-6)
    # +x pixels look toward +x world (flip=false: red wall x=-1 on the left), +y pixels look down
    _, dl = cbox_oracle.camera_generate(0.0, 256.0)
    _, dt = cbox_oracle.camera_generate(256.0, 0.0)
    assert dl[0] < 0 and abs(dl[1]) < 1e-6 and dt[1] > 0
    # Fov::Y(v) is used as v*aspect degrees across the image height (camera.rs:41-44)
    assert math.degrees(2 * math.atan2(dt[1], -dt[2])) == pytest.approx(19.5, rel=1e-4)


def test_fov_y_quirk_widescreen():
    sc = load_cbox(1920, 1080)
    osc = ob.OracleScene(sc)
    _, dt = osc.camera_generate(960.0, 0.0)
    half = math.degrees(math.atan2(dt[1], -dt[2]))
    # perspective(fov*aspect, 1) then the y scale by aspect: vertical half-angle = atan(tan(fov*a/2)/a)
    a = 1920 / 1080
    assert half == pytest.approx(math.degrees(math.atan(math.tan(math.radians(19.5 * a / 2)) / a)), rel=1e-4)


# ---- emitters: emitter.rs:570-688,1566-1647; scene.rs:53-123 ---------------------------------------------
def test_sample_light_cbox(cbox_oracle):
    x = np.float32([0.1, 0.5, 0.2])
    rng = np.random.default_rng(5)
    area = 0.47 * 0.38
    for r, u0, u1 in rng.random((200, 3)):
        rec = cbox_oracle.sample_light(x, 0.3, float(r), float(u0), float(u1))
        assert rec["mesh"] == 7
        p = rec["p"]
        assert -0.24 - 1e-6 <= p[0] <= 0.23 + 1e-6 and p[1] == pytest.approx(1.98) and -0.22 - 1e-6 <= p[2] <= 0.16 + 1e-6
        assert np.allclose(rec["n"], [0, -1, 0], atol=1e-6)  # aligned with the shading normals (geometry.rs:308-312)
        dvec = p.astype(np.float64) - x
        dist = np.linalg.norm(dvec)
        geom = max(dvec[1] / dist, 0) / dist**2  # n.(-d) with n=(0,-1,0)
        assert rec["pdf"] == pytest.approx((1 / area) / geom, rel=2e-5)
        assert np.allclose(rec["weight"], np.float32([17, 12, 4]) * geom * area, rtol=2e-5)
        # direct_pdf for the same geometry equals the sampling pdf
        assert cbox_oracle.direct_pdf(7, x, p, rec["n"], rec["d"]) == pytest.approx(rec["pdf"], rel=1e-5)


def test_sample_light_back_facing_is_invalid(cbox_oracle):
    rec = cbox_oracle.sample_light(np.float32([0.0, 1.99, 0.0]), 0.5, 0.5, 0.3, 0.3)  # above the light
    assert rec["pdf"] == 0.0 and not rec["weight"].any()


def test_two_emitters_are_flux_weighted():
    sc = SceneLoaderManager().load_string(json.dumps({
        "camera": {"width": 8, "height": 8, "fov": 40, "to_world": [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, -1, 0, 0, 0, 3, 1]},
        "meshes": [
            {"material": {"type": "diffuse", "kd": [0, 0, 0]}, "emission": [1, 1, 1], "indices": [0, 1, 2],
             "P": [0, 1, 0, 1, 1, 0, 0, 1, 1]},
            {"material": {"type": "diffuse", "kd": [0, 0, 0]}, "emission": [3, 1, 1], "indices": [0, 1, 2],
             "P": [0, 2, 0, 1, 2, 0, 0, 2, 1]}]}), "json")
    osc = ob.OracleScene(sc)
    picks = [osc.sample_light(np.float32([0.2, 0, 0.2]), float(r), 0.5, 0.3, 0.3)["mesh"] for r in np.linspace(0, 0.999, 400)]
    assert np.mean(np.array(picks) == 1) == pytest.approx(0.75, abs=0.01)  # channel_max(flux): 1 vs 3


# ---- accelerators: accel.rs ---------------------------------------------------------------------------------
def test_bvh_matches_reference_shape_invariants(cbox_oracle):
    info = cbox_oracle.bvh_info()
    assert info["nprims"] == 36 and info["nnodes"] % 2 == 1 and info["nnodes"] <= 71
    # root box: union of compute_aabb_tri boxes (flat extents padded by +-1e-4, geometry.rs:430-437)
    assert np.allclose(info["root_min"], [-1.0001, -1.00174846e-4, -1.0001], atol=1e-7)
    assert np.allclose(info["root_max"], [1.0001, 2.0001, 1.0], atol=1e-7)


def test_bvh_vs_naive_random_rays(cbox_oracle):
    rng = np.random.default_rng(6)
    n = 20000
    o = (rng.uniform(-0.99, 0.99, (n, 3)) + [0, 1, 0]).astype(np.float32)
    d = rng.normal(size=(n, 3))
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    pb, tb = cbox_oracle.trace(o, d, ob.ACCEL_BVH)
    pn, tn = cbox_oracle.trace(o, d, ob.ACCEL_NAIVE)
    # the two accelerators may only disagree on exact ties in t (random origins inside the tall box
    # see its bottom face and the coplanar floor at the same distance); BVH order is then arbitrary
    diff = pb != pn
    assert diff.sum() < 20 and np.array_equal(tb[:, 0], tn[:, 0]) and np.array_equal(tb[~diff], tn[~diff])
    p1 = (rng.uniform(-0.99, 0.99, (n, 3)) + [0, 1, 0]).astype(np.float32)
    assert np.array_equal(cbox_oracle.visible(o, p1, ob.ACCEL_BVH), cbox_oracle.visible(o, p1, ob.ACCEL_NAIVE))


def test_primary_grid_bvh_vs_naive_differ_only_on_exact_ties(cbox_oracle):
    """Some pixel centres look exactly at the seams where two walls of the symmetric box meet; both
    triangles accept the ray at the same t and the reference's answer depends on traversal order.
    Brute-force order (lowest (mesh,tri)) is the canonical answer used for GPU parity."""
    pb, tb = cbox_oracle.primary_hits(ob.ACCEL_BVH)
    pn, tn = cbox_oracle.primary_hits(ob.ACCEL_NAIVE)
    diff = pb != pn
    assert diff.sum() < 200
    assert np.array_equal(tb[..., 0][diff], tn[..., 0][diff])          # same distance
    assert (pn[diff] < pb[diff]).all()                                  # naive keeps the lowest index
    assert np.array_equal(tb[~diff], tn[~diff])


def test_visible_semantics(cbox_oracle):
    # light centre from the floor: visible; through the tall box: occluded
    assert cbox_oracle.visible([[0.0, 0.01, 0.9]], [[0.0, 1.98, 0.0]])[0] == 1
    assert cbox_oracle.visible([[-0.4, 0.01, -0.3]], [[-0.4, 1.5, -0.3]])[0] == 0
    # segments shorter than tnear=1e-4 fail BVHAccel's root test and count as occluded (accel.rs:338-340)
    assert cbox_oracle.visible([[0, 1, 0]], [[0, 1, 5e-5]])[0] == 0


# ---- estimators ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kw", [dict(), dict(strategy=_abi.RL_STRATEGY_BSDF), dict(strategy=_abi.RL_STRATEGY_EMITTER),
                                dict(max_depth=3), dict(min_depth=2, max_depth=6), dict(rr_depth=3, max_depth=8),
                                dict(rr_depth=None), dict(single_scattering=True)])
def test_stream_estimator_equals_graph(cbox64, kw):
    osc = ob.OracleScene(cbox64)
    integ = _abi.path_desc(**kw)
    g, sg = osc.render(integ, 8, seed=3, cfg=ob.config(estimator=ob.EST_GRAPH))
    s, ss = osc.render(integ, 8, seed=3, cfg=ob.config(estimator=ob.EST_STREAM))
    assert (sg.segments, sg.shadow_rays, sg.hits, sg.shadow_visible) == (ss.segments, ss.shadow_rays, ss.hits, ss.shadow_visible)
    assert rel_l2(s, g) < 2e-7


def test_math_modes_agree(cbox64):
    osc = ob.OracleScene(cbox64)
    integ = _abi.path_desc()
    a, sa = osc.render(integ, 16, cfg=ob.config(math_mode=ob.MATH_LIBM))
    b, sb = osc.render(integ, 16, cfg=ob.config(math_mode=ob.MATH_SPEC))
    assert sa.segments == sb.segments
    assert rel_l2(a, b) < 1e-4


def test_sampler_modes_agree_statistically(cbox64):
    """Mode A (per-block xoshiro stream, the reference's) vs mode B (counter stream, the GPU's)."""
    osc = ob.OracleScene(cbox64)
    integ = _abi.path_desc()
    spp = 256
    a, _ = osc.render(integ, spp, sampler_mode=_abi.RL_SAMPLER_BLOCK_STREAM)
    a2, _ = osc.render(integ, spp, seed=1, sampler_mode=_abi.RL_SAMPLER_BLOCK_STREAM)
    b, _ = osc.render(integ, spp, sampler_mode=_abi.RL_SAMPLER_COUNTER)
    noise = rel_l2(a, a2)            # two independent mode-A renders
    assert rel_l2(a, b) < 1.3 * noise  # A-vs-B differs by Monte-Carlo noise only
    assert a.mean() == pytest.approx(b.mean(), rel=0.02)
    c, _ = osc.render(integ, spp, sampler_mode=_abi.RL_SAMPLER_BLOCK_STREAM, cfg=ob.config(seeding=ob.SEED_SPLITMIX64))
    assert a.mean() == pytest.approx(c.mean(), rel=0.02)


def test_strategies_are_unbiased_against_each_other(cbox64):
    osc = ob.OracleScene(cbox64)
    means = [osc.render(_abi.path_desc(strategy=s), 256, seed=s)[0].mean(axis=(0, 1)) for s in (0, 1, 2)]
    assert np.allclose(means[0], means[1], rtol=0.06) and np.allclose(means[0], means[2], rtol=0.03)


def test_render_rejects_what_the_reference_panics_on(cbox64):
    osc = ob.OracleScene(cbox64)
    with pytest.raises(ValueError):
        osc.render(_abi.path_desc(), 0)             # assert_ne!(nb_samples, 0), integrators/mod.rs:410
    with pytest.raises(ValueError):
        osc.render(_abi.path_desc(max_depth=1), 1)  # unwrap of the missing sensor edge, path.rs:154


def test_tile_partition_is_exact(cbox64):
    osc = ob.OracleScene(cbox64)
    integ = _abi.path_desc()
    full, sf = osc.render(integ, 4)
    parts = [osc.render(integ, 4, cfg=ob.config(rank=r, nranks=3)) for r in range(3)]
    assert np.array_equal(sum(p[0] for p in parts), full)
    assert sum(p[1].segments for p in parts) == sf.segments and sum(p[1].samples for p in parts) == sf.samples


# ---- closed-form radiance ----------------------------------------------------------------------------------------
def _floor_under_light_scene(res=4):
    # 2x2 floor at y=0, unit square light at y=1 facing down, camera looking straight down at the origin
    return json.dumps({
        "camera": {"width": res, "height": res, "fov": 1.0, "fov_axis": "y",
                   "to_world": [1, 0, 0, 0, 0, 0, -1, 0, 0, -1, 0, 0, 0, 0.5, 0, 1]},
        "meshes": [
            {"name": "floor", "material": {"type": "diffuse", "kd": [0.5, 0.5, 0.5]}, "indices": [0, 1, 2, 0, 2, 3],
             "P": [-1, 0, -1, -1, 0, 1, 1, 0, 1, 1, 0, -1], "N": [0, 1, 0] * 4},
            {"name": "light", "material": {"type": "diffuse", "kd": [0, 0, 0]}, "emission": [2, 2, 2],
             "indices": [0, 1, 2, 0, 2, 3], "P": [-0.5, 1, -0.5, 0.5, 1, -0.5, 0.5, 1, 0.5, -0.5, 1, 0.5], "N": [0, -1, 0] * 4}]})


def _form_factor_point_to_parallel_square(h, a):
    # differential area under the centre of a parallel square of side a at height h: 4x the corner formula
    x = y = (a / 2) / h

    def corner(X, Y):
        return (1 / (2 * math.pi)) * (X / math.sqrt(1 + X * X) * math.atan(Y / math.sqrt(1 + X * X)) +
                                      Y / math.sqrt(1 + Y * Y) * math.atan(X / math.sqrt(1 + Y * Y)))
    return 4 * corner(x, y)


@pytest.mark.parametrize("integ", [_abi.direct_desc(1, 1), _abi.direct_desc(2, 0), _abi.direct_desc(0, 2),
                                   _abi.path_desc(max_depth=3)])
def test_direct_lighting_matches_lambert_form_factor(integ):
    """Radiance of a diffuse floor point under a square lamp: L = rho/pi * Le * pi * F = rho * Le * F."""
    sc = SceneLoaderManager().load_string(_floor_under_light_scene(), "json")
    osc = ob.OracleScene(sc)
    # the camera sits below the lamp and looks down: it sees the floor centre (narrow 1 degree fov)
    img, _ = osc.render(integ, 4096, seed=9)
    expect = 0.5 * 2.0 * _form_factor_point_to_parallel_square(1.0, 1.0)
    assert img.mean() == pytest.approx(expect, rel=0.02)


def test_white_furnace():
    """Closed cube, every wall emits Le and reflects rho: L = Le / (1 - rho)."""
    rho, le = 0.5, 1.0
    faces = []
    c = [(-1, -1, -1), (1, -1, -1), (1, 1, -1), (-1, 1, -1), (-1, -1, 1), (1, -1, 1), (1, 1, 1), (-1, 1, 1)]
    quads = [((0, 1, 2, 3), (0, 0, 1)), ((5, 4, 7, 6), (0, 0, -1)), ((4, 0, 3, 7), (1, 0, 0)), ((1, 5, 6, 2), (-1, 0, 0)),
             ((4, 5, 1, 0), (0, 1, 0)), ((3, 2, 6, 7), (0, -1, 0))]
    for q, n in quads:
        P = [x for i in q for x in c[i]]
        faces.append({"material": {"type": "diffuse", "kd": [rho] * 3}, "emission": [le] * 3, "indices": [0, 1, 2, 0, 2, 3],
                      "P": P, "N": list(n) * 4})
    sc = SceneLoaderManager().load_string(json.dumps({
        "camera": {"width": 8, "height": 8, "fov": 60, "to_world": [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, -1, 0, 0, 0, 0.5, 1]},
        "meshes": faces}), "json")
    osc = ob.OracleScene(sc)
    for strat in (_abi.RL_STRATEGY_ALL, _abi.RL_STRATEGY_BSDF):
        img, _ = osc.render(_abi.path_desc(strategy=strat), 2048, seed=11)
        assert img.mean() == pytest.approx(le / (1 - rho), rel=0.02)


def test_soup_scene_bvh_vs_naive():
    sc = SceneLoaderManager().load_string(soup_scene(600, seed=3), "json")
    osc = ob.OracleScene(sc)
    rng = np.random.default_rng(8)
    n = 5000
    o = rng.uniform(-1, 1, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3))
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    pb, tb = osc.trace(o, d, ob.ACCEL_BVH)
    pn, tn = osc.trace(o, d, ob.ACCEL_NAIVE)
    assert (pb != pn).mean() < 1e-3 and np.array_equal(tb[pb == pn], tn[pb == pn])


from hypothesis import given, settings, strategies as st_  # noqa: E402


@settings(max_examples=12, deadline=None)
@given(st_.integers(0, 10_000), st_.sampled_from([50, 200, 800]))
def test_bvh_and_naive_agree_on_random_scenes(seed, ntris):
    """The oracle's two accelerators (BVHAccel with the SAH tree, accel.rs:101-344; NaiveAcceleration, accel.rs:14-77) on random
    triangle soups, not only on the Cornell box: the BVH never reports a NEARER hit than the brute-force loop (it can only
    cull), wherever both hit the same triangle (t, u, v) agree bit for bit, and they disagree on at most a few rays in 10^4 --
    exact ties and hits on a box face that BVHAccel's slab test loses by an ulp (what the device's tie / rim machinery reproduces)."""
    from conftest import soup_scene
    from rustlight_b200 import SceneLoaderManager
    osc = ob.OracleScene(SceneLoaderManager().load_string(soup_scene(ntris, seed), "json"))
    rng = np.random.default_rng(seed + 1)
    n = 4000
    o = rng.uniform(-1.2, 1.2, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3))
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    pb, tb = osc.trace(o, d, ob.ACCEL_BVH)
    pn, tn = osc.trace(o, d, ob.ACCEL_NAIVE)
    miss = 0xFFFFFFFF
    both = (pb != miss) & (pn != miss)
    assert not ((pb != miss) & (pn == miss)).any()          # a BVH hit is an accepted triangle: the brute-force loop sees it too
    assert (tb[both, 0] >= tn[both, 0]).all()               # ... and finds nothing farther
    same = both & (pb == pn)
    assert np.array_equal(tb[same], tn[same])
    assert (pb != pn).sum() <= max(2, n // 2000)
