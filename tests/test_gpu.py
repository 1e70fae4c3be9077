"""GPU parity tests: everything goes through the C ABI (librl_b200.so) and is compared with the
oracle on the same seeded inputs, with the committed golden fixtures, and -- at BASELINE.json's
full sizes -- through size-independent properties.  Integer/index results must be identical;
radiance must be bit-identical to the oracle's stream estimator and within 1e-3 relative L2
(north_star tolerance) of the reference-faithful configuration."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import GOLDEN, author_metrics, load_cbox, rel_l2, soup_scene
from oracle import binding as ob
from rustlight_b200 import SceneLoaderManager, _abi
from rustlight_b200.device import Context, DeviceError, DeviceScene, IndependentSampler, IntegratorPathTracing, lib
from rustlight_b200.host import material_phong

pytestmark = pytest.mark.gpu

STREAM = dict(estimator=ob.EST_STREAM, accel_mode=ob.ACCEL_BVH)
TOL = 1e-3  # north_star: image within 1e-3 relative L2 of the CPU reference at matched spp


def _rays(n, seed, lo=-0.99, hi=0.99, shift=(0, 1, 0)):
    rng = np.random.default_rng(seed)
    o = (rng.uniform(lo, hi, (n, 3)) + shift).astype(np.float32)
    d = rng.normal(size=(n, 3))
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    p1 = (rng.uniform(lo, hi, (n, 3)) + shift).astype(np.float32)
    return o, d, p1


@pytest.fixture(scope="module")
def cbox_dev(gpu_ctx, cbox):
    d = DeviceScene(gpu_ctx, cbox)
    yield d
    d.close()


# ---- traversal kernel: exact hit indices ------------------------------------------------------------
def test_native_library_is_the_one_running(gpu_ctx, cbox_dev):
    assert lib()._name.endswith("rustlight_b200/librl_b200.so")
    bi = cbox_dev.bvh_info()
    # tree with leaves of <= 2 triangles (coherent rays) + the group table (incoherent rays: 18 planar quads -> 18 pair
    # records in 9 groups), all in shared memory
    assert (bi.ntris, bi.smem_resident) == (36, 1) and 18 <= bi.nleaves <= 36 and bi.nnodes == bi.nleaves - 1 and bi.max_depth <= 36
    assert (bi.flat_groups, bi.flat_pairs, bi.flat_singles) == (9, 18, 0) and bi.flat_delta < 1e-6


def test_primary_ray_grid_exact(cbox_dev, cbox_oracle):
    """north_star: exact match on ray/triangle hit indices for the fixed primary-ray test -- against the reference's DEFAULT
    accelerator (BVHAccel: ties at wall seams go to whichever leaf it visits first), on all 262 144 pixels, (t, u, v) included."""
    pg, tg = cbox_dev.primary_hits()
    po, to = cbox_oracle.primary_hits(ob.ACCEL_BVH)
    assert np.array_equal(pg, po) and np.array_equal(tg, to)
    g = np.load(os.path.join(GOLDEN, "cbox512_primary_hits.npz"))
    assert np.array_equal(np.where(pg == 0xFFFFFFFF, 255, pg).astype(np.uint8), g["prim"])
    assert np.array_equal(tg[::8, ::8, 0], g["t_sub8"])
    # the brute-force accelerator of the reference (accel.rs:22-51) answers differently on the exact ties: the device follows BVHAccel there
    pn, tn = cbox_oracle.primary_hits(ob.ACCEL_NAIVE)
    assert 0 < (pg != pn).sum() < 200 and np.array_equal(tg[..., 0], tn[..., 0])


def test_random_rays_exact(cbox_dev, cbox_oracle):
    o, d, p1 = _rays(300000, 1)
    pg, tg = cbox_dev.trace(o, d)
    po, to = cbox_oracle.trace(o, d, ob.ACCEL_BVH)
    assert np.array_equal(pg, po) and np.array_equal(tg, to)
    assert np.array_equal(cbox_dev.visible(o, p1), cbox_oracle.visible(o, p1, ob.ACCEL_BVH))


def test_edge_case_rays(cbox_dev, cbox_oracle):
    rng = np.random.default_rng(3)
    o = (rng.uniform(-0.9, 0.9, (5000, 3)) + [0, 1, 0]).astype(np.float32)
    axes = np.eye(3, dtype=np.float32)[rng.integers(0, 3, 5000)] * rng.choice([-1, 1], (5000, 1)).astype(np.float32)  # +-0 components
    tgt = (rng.uniform(-1, 1, (5000, 3)) + [0, 1, 0])
    far = (tgt + rng.normal(size=(5000, 3)) * 300).astype(np.float32)
    dfar = tgt - far
    dfar = (dfar / np.linalg.norm(dfar, axis=1, keepdims=True)).astype(np.float32)
    for oo, dd in ((o, axes), (far, dfar)):
        pg, tg = cbox_dev.trace(oo, dd)
        po, to = cbox_oracle.trace(oo, dd, ob.ACCEL_BVH)
        assert np.array_equal(pg, po) and np.array_equal(tg, to)
    # empty input, zero-length segment, segment shorter than tnear
    assert cbox_dev.trace(np.zeros((0, 3)), np.zeros((0, 3)))[0].size == 0
    p = np.float32([[0, 1, 0], [0, 1, 0]])
    q = np.float32([[0, 1, 0], [0, 1, 5e-5]])
    assert np.array_equal(cbox_dev.visible(p, q), cbox_oracle.visible(p, q, ob.ACCEL_BVH))


@pytest.mark.parametrize("ntris,seed", [(1, 0), (2, 1), (50, 2), (600, 3), (4000, 4)])
def test_soup_scenes_exact(gpu_ctx, ntris, seed):
    """Ragged sizes incl. a single triangle; 4000 triangles do not fit shared memory (global-memory path)."""
    if ntris <= 2:
        import json
        tris = [[0, 0, 0, 1, 0, 0, 0, 1, 0], [0.2, 0.2, -0.5, 1.2, 0.2, -0.5, 0.2, 1.2, -0.5]][:ntris]
        meshes = [{"material": {"type": "diffuse", "kd": [0.5] * 3}, "emission": [1, 1, 1], "indices": [0, 1, 2], "P": t} for t in tris]
        txt = json.dumps({"camera": {"width": 16, "height": 16, "fov": 40, "to_world": [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, -1, 0, 0.3, 0.3, 3, 1]}, "meshes": meshes})
    else:
        txt = soup_scene(ntris, seed)
    sc = SceneLoaderManager().load_string(txt, "json")
    dev, osc = DeviceScene(gpu_ctx, sc), ob.OracleScene(sc)
    bi = dev.bvh_info()
    assert bi.ntris == sc.nb_triangles and bi.smem_resident == (1 if ntris < 200 else 0)
    o, d, p1 = _rays(50000, seed + 20, -1.2, 1.2, (0, 0, 0))
    pg, tg = dev.trace(o, d)
    po, to = osc.trace(o, d, ob.ACCEL_BVH)
    assert np.array_equal(pg, po) and np.array_equal(tg, to)
    assert np.array_equal(dev.visible(o, p1), osc.visible(o, p1, ob.ACCEL_BVH))
    pg, tg = dev.primary_hits()
    po, to = osc.primary_hits(ob.ACCEL_BVH)
    assert np.array_equal(pg, po) and np.array_equal(tg, to)
    dev.close()


# ---- the wavefront: radiance parity ---------------------------------------------------------------------
@pytest.mark.parametrize("kw", [dict(), dict(strategy=_abi.RL_STRATEGY_BSDF), dict(strategy=_abi.RL_STRATEGY_EMITTER),
                                dict(max_depth=2), dict(max_depth=4, min_depth=2), dict(rr_depth=4, max_depth=9),
                                dict(rr_depth=None), dict(single_scattering=True)])
def test_path_render_bit_exact_vs_oracle(gpu_ctx, kw):
    sc = load_cbox(200, 136)  # ragged: neither dimension is a multiple of the 16x16 tile
    dev = DeviceScene(gpu_ctx, sc)
    integ = _abi.path_desc(**kw)
    img, st = dev.render(integ, 6, seed=5, batch_spp=4)  # 2 batches, the second one ragged
    ref, so = ob.OracleScene(sc).render(integ, 6, seed=5, cfg=ob.config(**STREAM))
    assert (st.samples, st.segments, st.hits, st.shadow_rays, st.shadow_visible) == (so.samples, so.segments, so.hits, so.shadow_rays, so.nee_added)
    assert np.array_equal(img, ref)
    dev.close()


def test_cbox_512_spp16_vs_faithful_oracle(cbox_dev, cbox_oracle):
    """BASELINE configs[0]: path -n 16, 512x512.  GPU vs the reference-faithful oracle configuration
    (graph estimator, BVHAccel order, glibc math) on the same counter-based stream."""
    integ = _abi.path_desc()
    img, st = cbox_dev.render(integ, 16, seed=0)
    ref, so = cbox_oracle.render(integ, 16, seed=0, cfg=ob.config(math_mode=ob.MATH_LIBM, accel_mode=ob.ACCEL_BVH, estimator=ob.EST_GRAPH))
    r = rel_l2(img, ref)
    assert r < TOL, r
    m = author_metrics(ref, img)  # the reference author's own metrics (metric.py:18-45, eps = 1e-2)
    print("C1 GPU vs faithful oracle: rel_l2 %.3g, " % r + ", ".join("%s %.3g" % kv for kv in m.items()))
    assert m["mape"] < TOL and m["smape"] < TOL and m["l1"] < TOL and m["mrse"] < TOL * TOL
    assert abs(int(st.segments) - int(so.segments)) <= 32 and abs(int(st.shadow_rays) - int(so.shadow_rays)) <= 32
    ex, se = cbox_oracle.render(integ, 16, seed=0, cfg=ob.config(**STREAM))
    assert np.array_equal(img, ex) and st.segments == se.segments


def test_golden_fixture(gpu_ctx):
    g = np.load(os.path.join(GOLDEN, "cbox64_spp16_seed0.npz"))
    dev = DeviceScene(gpu_ctx, load_cbox(64, 64))
    for name, integ in [("path", _abi.path_desc()), ("path_bsdf", _abi.path_desc(strategy=_abi.RL_STRATEGY_BSDF)), ("path_d4", _abi.path_desc(max_depth=4))]:
        img, st = dev.render(integ, 16, seed=0)
        assert rel_l2(img, g[name]) < TOL
        assert abs(int(st.segments) - int(g[name + "_counts"][0])) <= 4
    dev.close()


def test_phong_walls_bit_exact(gpu_ctx):
    """BASELINE configs[2]: Cornell box with glossy Phong walls (kd = 0.5*wall colour, ks = 0.3, exponent 50)."""
    sc = load_cbox(128, 128)
    kds = [(0.63, 0.065, 0.05), (0.14, 0.45, 0.091), (0.725, 0.71, 0.68)]
    for mesh, kd in [(0, kds[2]), (1, kds[2]), (2, kds[2]), (3, kds[1]), (4, kds[0])]:
        sc.set_material(mesh, material_phong([0.5 * c for c in kd], (0.3, 0.3, 0.3), 50.0))
    dev = DeviceScene(gpu_ctx, sc)
    integ = _abi.path_desc()
    img, st = dev.render(integ, 8, seed=2)
    ref, so = ob.OracleScene(sc).render(integ, 8, seed=2, cfg=ob.config(**STREAM))
    assert (st.segments, st.hits, st.shadow_rays) == (so.segments, so.hits, so.shadow_rays)
    assert np.array_equal(img, ref)
    faithful, _ = ob.OracleScene(sc).render(integ, 8, seed=2, cfg=ob.config(math_mode=ob.MATH_LIBM))
    assert rel_l2(img, faithful) < TOL
    dev.close()


def test_soup_render_bit_exact(gpu_ctx):
    sc = SceneLoaderManager().load_string(soup_scene(1500, 4, 64, 64), "json")
    dev = DeviceScene(gpu_ctx, sc)
    integ = _abi.path_desc(max_depth=6)
    img, st = dev.render(integ, 4, seed=1)
    ref, so = ob.OracleScene(sc).render(integ, 4, seed=1, cfg=ob.config(**STREAM))
    assert (st.segments, st.hits, st.shadow_rays) == (so.segments, so.hits, so.shadow_rays)
    assert np.array_equal(img, ref)
    dev.close()


# ---- properties at BASELINE.json's full size (configs[1]: 1024x1024, 128 spp) ---------------------------------
@pytest.fixture(scope="module")
def c2(gpu_ctx):
    sc = load_cbox().scale_image(2.0)  # CLI `-s 2`: 1024x1024, matrices untouched
    dev = DeviceScene(gpu_ctx, sc)
    yield sc, dev
    dev.close()


def test_full_size_deterministic_and_batch_independent(c2):
    sc, dev = c2
    integ = _abi.path_desc()
    a, sa = dev.render(integ, 128, seed=0)
    b, sb = dev.render(integ, 128, seed=0, batch_spp=5)  # different wavefront batching, ragged last batch
    assert np.array_equal(a, b) and sa.segments == sb.segments and sa.shadow_visible == sb.shadow_visible
    assert sa.samples == 1024 * 1024 * 128 and sa.hits <= sa.segments and sa.shadow_visible <= sa.shadow_rays == sa.hits
    assert np.isfinite(a).all() and (a >= 0).all()
    # the same pixels at 1/64 of the work agree within Monte-Carlo noise; the estimate is consistent
    c, _ = dev.render(integ, 2, seed=7)
    assert a.mean(axis=(0, 1)) == pytest.approx(c.mean(axis=(0, 1)), rel=0.01)
    # against the oracle on a sub-sample of rows (bit-exact, same stream)
    osc = ob.OracleScene(sc)
    for px, py in [(5, 7), (512, 512), (1000, 30), (300, 900)]:
        acc = np.zeros(3, np.float32)
        for s in range(128):
            rgb, *_ = osc.path_sample(integ, 0, px, py, s, ob.config(**STREAM))
            acc = (acc + rgb).astype(np.float32)
        assert np.array_equal(acc * np.float32(1.0 / 128.0), a[py, px])


def test_full_size_strategies_are_unbiased(c2):
    _, dev = c2
    m = [dev.render(_abi.path_desc(strategy=s), 8, seed=s)[0].mean(axis=(0, 1)) for s in (0, 1, 2)]
    assert np.allclose(m[0], m[1], rtol=0.01) and np.allclose(m[0], m[2], rtol=0.01)


def test_tile_partition_sums_to_the_full_image(cbox):
    """Multi-GPU decomposition on one device: rank images are disjoint and add up exactly."""
    integ = _abi.path_desc()
    full_ctx = Context(0)
    full, sf = DeviceScene(full_ctx, cbox).render(integ, 4, seed=9)
    parts, segs = [], 0
    for r in range(3):
        ctx = Context(0, nranks=3, rank=r)
        img, st = DeviceScene(ctx, cbox).render(integ, 4, seed=9)
        parts.append(img)
        segs += st.segments
        ctx.close()
    assert np.array_equal(parts[0] + parts[1] + parts[2], full) and segs == sf.segments
    assert not (parts[0].astype(bool) & parts[1].astype(bool)).any()
    full_ctx.close()


def test_integrator_mirror_api(gpu_ctx, cbox64):
    cbox64.nb_samples = 4
    bc = IntegratorPathTracing().compute(IndependentSampler(3), DeviceScene(gpu_ctx, cbox64))
    ref, _ = ob.OracleScene(cbox64).render(_abi.path_desc(), 4, seed=3, cfg=ob.config(**STREAM))
    assert np.array_equal(bc.values["primal"], ref)


# ---- error behaviour (the reference panics; the C ABI returns codes) ---------------------------------------------
def test_error_codes(gpu_ctx, cbox_dev, cbox):
    with pytest.raises(DeviceError) as e:
        cbox_dev.render(_abi.path_desc(), 0)                 # assert_ne!(nb_samples, 0)
    assert e.value.code == _abi.RL_ERR_INVALID
    with pytest.raises(DeviceError) as e:
        cbox_dev.render(_abi.path_desc(max_depth=1), 1)      # path.rs:154 unwrap
    assert e.value.code == _abi.RL_ERR_INVALID
    desc = cbox.desc.contents
    bad = _abi.rl_scene_desc(desc.nmeshes, desc.meshes, desc.camera, 1, 0)  # scene.volume = Some(..)
    h = C.c_void_p()
    assert lib().rl_scene_create(gpu_ctx._h, C.byref(bad), C.byref(h)) == _abi.RL_ERR_UNSUPPORTED
    assert b"volume" in lib().rl_last_error(gpu_ctx._h)
    opts = _abi.rl_render_opts(C.sizeof(_abi.rl_render_opts), 1, 0, _abi.RL_SAMPLER_BLOCK_STREAM, 0, 0, 0)
    st = _abi.rl_stats()
    integ = _abi.path_desc()
    assert lib().rl_render(gpu_ctx._h, cbox_dev._h, C.byref(integ), C.byref(opts), None, C.byref(st)) == _abi.RL_ERR_UNSUPPORTED


# ---- `direct` integrator (BASELINE configs[3]) ----------------------------------------------------------------------
@pytest.mark.parametrize("nb,nl", [(1, 1), (2, 3), (0, 2), (3, 0)])
def test_direct_render_bit_exact_vs_oracle(gpu_ctx, nb, nl):
    sc = load_cbox(200, 136)
    dev = DeviceScene(gpu_ctx, sc)
    integ = _abi.direct_desc(nb, nl)
    img, st = dev.render(integ, 6, seed=4, batch_spp=4)
    ref, so = ob.OracleScene(sc).render(integ, 6, seed=4, cfg=ob.config(accel_mode=ob.ACCEL_BVH))
    assert (st.samples, st.segments, st.hits, st.shadow_rays) == (so.samples, so.segments, so.hits, so.shadow_rays)
    assert np.array_equal(img, ref)
    faithful, _ = ob.OracleScene(sc).render(integ, 6, seed=4, cfg=ob.config(math_mode=ob.MATH_LIBM, accel_mode=ob.ACCEL_BVH))
    assert rel_l2(img, faithful) < TOL
    dev.close()


def test_direct_golden_and_full_size_property(gpu_ctx):
    g = np.load(os.path.join(GOLDEN, "cbox64_spp16_seed0.npz"))
    dev = DeviceScene(gpu_ctx, load_cbox(64, 64))
    img, st = dev.render(_abi.direct_desc(1, 1), 16, seed=0)
    assert rel_l2(img, g["direct"]) < TOL and abs(int(st.segments) - int(g["direct_counts"][0])) <= 4
    dev.close()
    # configs[3] size: 2048x2048 (`-s 4`), direct -n 64 is 3 rays per sample; here 4 spp of it
    big = DeviceScene(gpu_ctx, load_cbox().scale_image(4.0))
    a, sa = big.render(_abi.direct_desc(1, 1), 4, seed=1)
    b, sb = big.render(_abi.direct_desc(1, 1), 4, seed=1, batch_spp=1)
    assert np.array_equal(a, b) and sa.segments == sb.segments
    assert sa.samples == 2048 * 2048 * 4 and sa.segments <= 2 * sa.samples and sa.shadow_rays <= sa.samples
    # direct lighting is the depth-2 truncation of the path integrator: same expectation
    p, _ = big.render(_abi.path_desc(max_depth=3), 4, seed=2)
    assert a.mean(axis=(0, 1)) == pytest.approx(p.mean(axis=(0, 1)), rel=0.02)
    big.close()


# ---- material sort (BASELINE configs[4]: "with per-bounce material sort") -----------------------------------------------
def test_material_sort_changes_nothing_but_the_schedule(gpu_ctx):
    sc = load_cbox(200, 136)
    kds = [(0.63, 0.065, 0.05), (0.14, 0.45, 0.091), (0.725, 0.71, 0.68)]
    for mesh, kd in [(0, kds[2]), (2, kds[2]), (4, kds[0])]:  # floor, back wall, left wall become Phong: two BSDF kinds
        sc.set_material(mesh, material_phong([0.5 * c for c in kd], (0.3, 0.3, 0.3), 50.0))
    dev = DeviceScene(gpu_ctx, sc)
    integ = _abi.path_desc()
    a, sa = dev.render(integ, 6, seed=8)
    b, sb = dev.render(integ, 6, seed=8, material_sort=1)
    assert np.array_equal(a, b)
    assert (sa.segments, sa.hits, sa.shadow_rays, sa.shadow_visible) == (sb.segments, sb.hits, sb.shadow_rays, sb.shadow_visible)
    ref, so = ob.OracleScene(sc).render(integ, 6, seed=8, cfg=ob.config(**STREAM))
    assert np.array_equal(b, ref) and sb.segments == so.segments
    dev.close()



@pytest.mark.parametrize("seed", range(6))
def test_adversarial_scenes_exact(gpu_ctx, seed):
    """Same stress set as the CPU emulator test, on the device (approximate reciprocal in the prefilter)."""
    from conftest import adversarial_case
    sc, o, dd, p1 = adversarial_case(seed)
    dev, osc = DeviceScene(gpu_ctx, sc), ob.OracleScene(sc)
    pg, tg = dev.trace(o, dd)
    po, to = osc.trace(o, dd, ob.ACCEL_BVH)
    assert np.array_equal(pg, po) and np.array_equal(tg, to)
    assert np.array_equal(dev.visible(o, p1), osc.visible(o, p1, ob.ACCEL_BVH))
    dev.close()


@pytest.mark.parametrize("seed", range(4))
def test_adversarial_quads_exact(gpu_ctx, seed):
    """Pair records of the group table (packed f32x2 scan, sign-bit reject mask) on near-coplanar quads: rl_trace runs
    k_trace_flat, rl_visible runs flat_any."""
    from conftest import adversarial_pairs_case
    sc, o, dd, p1 = adversarial_pairs_case(seed)
    dev, osc = DeviceScene(gpu_ctx, sc), ob.OracleScene(sc)
    bi = dev.bvh_info()
    assert bi.flat_groups > 0 and bi.flat_pairs >= 3
    pg, tg = dev.trace(o, dd)
    po, to = osc.trace(o, dd, ob.ACCEL_BVH)
    assert np.array_equal(pg, po) and np.array_equal(tg, to)
    assert np.array_equal(dev.visible(o, p1), osc.visible(o, p1, ob.ACCEL_BVH))
    dev.close()


def test_group_table_on_surface_rays_and_degenerate_directions(cbox_dev, cbox_oracle):
    """What the wavefront actually traces: rays leaving the surfaces (origins ON planes of the table: num == 0 for the
    own quad), axis-parallel directions (d.n == 0: inf/NaN arithmetic in the scan must stay a candidate), segments to the light."""
    o, d, _ = _rays(300000, 21)
    po, to = cbox_oracle.trace(o, d, ob.ACCEL_BVH)
    hit = po != 0xFFFFFFFF
    o2 = (o[hit] + d[hit] * to[hit, :1]).astype(np.float32)
    d2 = _rays(len(o2), 22)[1]
    axes = np.eye(3, dtype=np.float32)[np.arange(len(o2)) % 3] * np.where(np.arange(len(o2)) % 2, 1, -1).astype(np.float32)[:, None]
    d2[::7] = axes[::7]
    pg, tg = cbox_dev.trace(o2, d2)
    po2, to2 = cbox_oracle.trace(o2, d2, ob.ACCEL_BVH)
    assert np.array_equal(pg, po2) and np.array_equal(tg, to2)
    light = (np.array([[0.0, 1.98, -0.03]]) + (_rays(len(o2), 23)[0] - [0, 1, 0]) * [0.24, 0, 0.2]).astype(np.float32)
    vg = cbox_dev.visible(o2, light)
    assert np.array_equal(vg, cbox_oracle.visible(o2, light, ob.ACCEL_BVH)) and 0.2 < vg.mean() < 0.9


@pytest.mark.parametrize("sort", [0, 1])
def test_mixed_material_scene_bit_exact(gpu_ctx, sort):
    """Every BSDF kind in one scene (metal GGX / Beckmann, mirror, glass, substrate with and without a distribution, phong,
    diffuse): PDF::Discrete edges, smooth vertices without light sampling, the non-two-sided glass box -- with and
    without the per-bounce material sort (7 keys)."""
    from conftest import mixed_cbox
    sc = mixed_cbox(96, 96)
    dev, osc = DeviceScene(gpu_ctx, sc), ob.OracleScene(sc)
    for kw in (dict(), dict(strategy=_abi.RL_STRATEGY_BSDF), dict(max_depth=6, rr_depth=2)):
        integ = _abi.path_desc(**kw)
        img, st = dev.render(integ, 8, seed=12, material_sort=sort)
        ref, so = osc.render(integ, 8, seed=12, cfg=ob.config(**STREAM))
        assert (st.segments, st.hits, st.shadow_rays) == (so.segments, so.hits, so.shadow_rays)
        assert np.array_equal(img, ref)
    if sort == 0:
        for nb, nl in ((1, 1), (2, 2)):
            integ = _abi.direct_desc(nb, nl)
            img, st = dev.render(integ, 4, seed=5)
            ref, so = osc.render(integ, 4, seed=5, cfg=ob.config(**STREAM))
            assert np.array_equal(img, ref) and st.segments == so.segments
    dev.close()


def test_light_tree_bit_exact(gpu_ctx):
    """`-x ats` (LightSamplerATS, emitter.rs:782-1400): 26 emissive triangles in 10 meshes, light sampling through the importance-driven
    descent of the light tree; `path` (tree with Some(n_s) for light samples, with None for the MIS pdf of BSDF-sampled hits), `direct`
    with light and BSDF samples."""
    from test_ats import many_lights_scene
    sc = many_lights_scene(96, 96)
    dev, osc = DeviceScene(gpu_ctx, sc), ob.OracleScene(sc)
    # (`direct` with BSDF samples: the tree's pdf of the second hit takes the FIRST vertex' shading normal, Some(&its.n_s), direct.rs:158-165)
    for integ in (_abi.path_desc(), _abi.path_desc(strategy=_abi.RL_STRATEGY_EMITTER), _abi.path_desc(max_depth=4, rr_depth=2), _abi.direct_desc(0, 2),
                  _abi.direct_desc(1, 1), _abi.direct_desc(2, 1)):
        img, st = dev.render(integ, 8, seed=12)
        ref, so = osc.render(integ, 8, seed=12, cfg=ob.config(**STREAM))
        assert (st.segments, st.hits, st.shadow_rays) == (so.segments, so.hits, so.shadow_rays)
        assert np.array_equal(img, ref)
    dev.close()
    box = load_cbox(64, 64).set_ats(True)  # the plain Cornell box: a tree of two leaves
    dev, osc = DeviceScene(gpu_ctx, box), ob.OracleScene(box)
    img, st = dev.render(_abi.path_desc(), 8, seed=2)
    ref, so = osc.render(_abi.path_desc(), 8, seed=2, cfg=ob.config(**STREAM))
    assert st.segments == so.segments and np.array_equal(img, ref)
    dev.close()


def test_varying_emission_bit_exact(gpu_ctx):
    """`-x hvs-light` / `-x texture-light` (EmissionType::HSV / Texture, geometry.rs:99-104, 184-206): the lamp's emission depends on the uv
    of the hit (arrival emission) and on the normalized uv of the sampled point (light sampling); `path` in all three strategies, `direct`,
    the light tree, the tail kernel and the tessellated box (tree kernels)."""
    from test_emission import hsv_box, texture_box
    for make in (hsv_box, texture_box):
        sc = make(96, 96)
        dev, osc = DeviceScene(gpu_ctx, sc), ob.OracleScene(sc)
        for integ in (_abi.path_desc(), _abi.path_desc(strategy=_abi.RL_STRATEGY_EMITTER), _abi.path_desc(strategy=_abi.RL_STRATEGY_BSDF, max_depth=4),
                      _abi.direct_desc(2, 2), _abi.direct_desc(0, 1)):
            img, st = dev.render(integ, 8, seed=21)
            ref, so = osc.render(integ, 8, seed=21, cfg=ob.config(**STREAM))
            assert (st.segments, st.hits, st.shadow_rays) == (so.segments, so.hits, so.shadow_rays)
            assert np.array_equal(img, ref)
        dev.close()
    sc = hsv_box(64, 64).set_ats(True)
    dev, osc = DeviceScene(gpu_ctx, sc), ob.OracleScene(sc)
    img, st = dev.render(_abi.path_desc(), 8, seed=5)
    ref, so = osc.render(_abi.path_desc(), 8, seed=5, cfg=ob.config(**STREAM))
    assert st.segments == so.segments and np.array_equal(img, ref)
    dev.close()
    # a lamp of one colour changes nothing: the kernels with the emission lookup equal the plain ones
    plain = load_cbox(64, 64)
    a, _ = DeviceScene(gpu_ctx, plain).render(_abi.path_desc(), 4, seed=9)
    const = load_cbox(64, 64)
    const.override_lights_texture(const.add_bitmap_texture(np.ones((2, 2, 3), np.float32)))
    b, _ = DeviceScene(gpu_ctx, const).render(_abi.path_desc(), 4, seed=9)
    lum = np.float32(17.0) * np.float32(0.212671) + np.float32(12.0) * np.float32(0.715160) + np.float32(4.0) * np.float32(0.072169)
    assert a.mean() > 0 and abs(b[..., 0].mean() / a[..., 0].mean() - lum / 17.0) < 0.05  # same picture up to the lamp's colour


def test_mitsuba_xml_scene_bit_exact(gpu_ctx):
    """The Mitsuba-XML route (the reference's only way to a BSDFPhong): Phong walls, checkerboard floor, rough plastic, a rough-conductor
    sphere of 1922 triangles (4-wide tree over the reference's topology), area + point light."""
    from test_mitsuba import BOX
    sc = SceneLoaderManager().load_string(BOX, "xml")
    sc.set_resolution(120, 96)
    dev, osc = DeviceScene(gpu_ctx, sc), ob.OracleScene(sc)
    assert dev.bvh_info().flat_groups == 0 and dev.bvh_info().ntris == 1934
    for integ in (_abi.path_desc(), _abi.path_desc(max_depth=5, rr_depth=2), _abi.direct_desc(2, 2)):
        img, st = dev.render(integ, 8, seed=7)
        ref, so = osc.render(integ, 8, seed=7, cfg=ob.config(**STREAM))
        assert (st.segments, st.hits, st.shadow_rays) == (so.segments, so.hits, so.shadow_rays)
        assert np.array_equal(img, ref)
    pg, tg = dev.primary_hits()
    po, to = osc.primary_hits(ob.ACCEL_BVH)
    assert np.array_equal(pg, po) and np.array_equal(tg, to)
    dev.close()


@pytest.mark.parametrize("tail", [None, "0"])
def test_environment_texture_bit_exact(gpu_ctx, tail, monkeypatch):
    """EnvironmentLightColor::Texture (emitter.rs:300-427): a lat-long HDR image with a sun texel, a black row and a black texel as the
    environment of an open scene (octahedron over a glossy floor + an area light) and of the Cornell box; `path` with every strategy,
    `direct`; with the tail kernel and as a pure wavefront."""
    from test_envmap import env_scene, sky_image
    if tail is not None:
        monkeypatch.setenv("RL_TAIL_MAX", tail)
    sc = env_scene(sky_image(), 96, 96, area_light=True, floor=True)
    dev, osc = DeviceScene(gpu_ctx, sc), ob.OracleScene(sc)
    for integ in (_abi.path_desc(), _abi.path_desc(strategy=_abi.RL_STRATEGY_BSDF), _abi.path_desc(strategy=_abi.RL_STRATEGY_EMITTER),
                  _abi.path_desc(max_depth=4, rr_depth=2), _abi.direct_desc(1, 1), _abi.direct_desc(0, 2), _abi.direct_desc(2, 0)):
        img, st = dev.render(integ, 8, seed=12)
        ref, so = osc.render(integ, 8, seed=12, cfg=ob.config(**STREAM))
        assert (st.segments, st.hits, st.shadow_rays) == (so.segments, so.hits, so.shadow_rays)
        assert np.array_equal(img, ref)
    dev.close()
    if tail is None:
        box = load_cbox(64, 64)
        box.set_environment_texture(box.add_bitmap_texture(sky_image(seed=3)))
        dev, osc = DeviceScene(gpu_ctx, box), ob.OracleScene(box)
        img, st = dev.render(_abi.path_desc(), 8, seed=2)
        ref, so = osc.render(_abi.path_desc(), 8, seed=2, cfg=ob.config(**STREAM))
        assert st.segments == so.segments and np.array_equal(img, ref)
        dev.close()


@pytest.mark.parametrize("sort", [0, 1])
def test_blended_materials_bit_exact(gpu_ctx, sort):
    """BSDFBlend (bsdfs/blend.rs) on four meshes of the Cornell box, every rough kind as a part: `path` (all strategies' MIS terms go
    through the blend's pdf / eval), `direct`, the tail kernel and the material sort (blend = its own key)."""
    from conftest import blended_cbox
    sc = blended_cbox(96, 96)
    dev, osc = DeviceScene(gpu_ctx, sc), ob.OracleScene(sc)
    for kw in (dict(), dict(strategy=_abi.RL_STRATEGY_BSDF), dict(max_depth=6, rr_depth=2)):
        integ = _abi.path_desc(**kw)
        img, st = dev.render(integ, 8, seed=12, material_sort=sort)
        ref, so = osc.render(integ, 8, seed=12, cfg=ob.config(**STREAM))
        assert (st.segments, st.hits, st.shadow_rays) == (so.segments, so.hits, so.shadow_rays)
        assert np.array_equal(img, ref)
    if sort == 0:
        integ = _abi.direct_desc(1, 2)
        img, st = dev.render(integ, 4, seed=5)
        ref, so = osc.render(integ, 4, seed=5, cfg=ob.config(**STREAM))
        assert np.array_equal(img, ref) and st.segments == so.segments
    dev.close()


@pytest.mark.parametrize("dist,nc", [(1.0, False), (None, True), (0.25, False)])
def test_ao_bit_exact(gpu_ctx, cbox, dist, nc):
    dev, osc = DeviceScene(gpu_ctx, cbox), ob.OracleScene(cbox)
    integ = _abi.ao_desc(dist, nc)
    img, st = dev.render(integ, 4, seed=6)
    ref, so = osc.render(integ, 4, seed=6, cfg=ob.config(**STREAM))
    assert st.segments == so.segments and np.array_equal(img, ref)
    dev.close()


def test_point_and_directional_lights_bit_exact(gpu_ctx):
    """PointEmitter / DirectionalLight next to the area light: flux-weighted selection over three emitters, PDF::Discrete
    light edges (no MIS), shadow segments that end outside the scene (directional)."""
    sc = load_cbox(96, 96)
    sc.add_point_light((0.6, 0.5, 0.4), (0.3, 1.2, 0.4))
    sc.add_directional_light((0.8, 0.8, 1.0), (0.3, -1.0, -0.2))
    dev, osc = DeviceScene(gpu_ctx, sc), ob.OracleScene(sc)
    for integ in (_abi.path_desc(), _abi.path_desc(strategy=_abi.RL_STRATEGY_EMITTER, max_depth=3), _abi.direct_desc(1, 2)):
        img, st = dev.render(integ, 6, seed=7)
        ref, so = osc.render(integ, 6, seed=7, cfg=ob.config(**STREAM))
        assert (st.segments, st.hits, st.shadow_rays) == (so.segments, so.hits, so.shadow_rays)
        assert np.array_equal(img, ref)
    dev.close()


@pytest.mark.parametrize("n", [2, 24])
def test_tessellated_cornell_box_through_the_lbvh(gpu_ctx, n):
    """The Cornell box with every face cut into n x n cells (tools/tess_cbox.py): 36 n^2 triangles, no group table, so
    every ray walks the LBVH (shared-memory resident for n = 2: 23 KB, global memory for n = 24).  Bit-exact against the oracle,
    and nearly the same picture as the plain box (same random streams: only paths through cell edges change)."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    from tess_cbox import tessellated_cbox_json
    sc = SceneLoaderManager().load_string(tessellated_cbox_json(n), "json")
    sc.set_resolution(64, 64)
    dev = DeviceScene(gpu_ctx, sc)
    bi = dev.bvh_info()
    assert bi.ntris == 36 * n * n and bi.flat_groups == 0 and bi.smem_resident == (1 if n == 2 else 0)
    img, st = dev.render(_abi.path_desc(), 8, seed=2)
    ref, so = ob.OracleScene(sc).render(_abi.path_desc(), 8, seed=2, cfg=ob.config(**STREAM))
    assert st.segments == so.segments and np.array_equal(img, ref)
    plain, _ = ob.OracleScene(load_cbox(64, 64)).render(_abi.path_desc(), 8, seed=2, cfg=ob.config(**STREAM))
    assert rel_l2(img, plain) < 0.05
    dev.close()


def test_textured_materials_bit_exact(gpu_ctx):
    """BSDFColor::{Checkerbord, Bitmap, Grid} on diffuse slots of the Cornell box: uv interpolation + lookups on the device."""
    from rustlight_b200.host import material_diffuse, material_substrate
    sc = load_cbox(96, 96)
    t1 = sc.add_checkerboard_texture((0.8, 0.8, 0.8), (0.1, 0.1, 0.1), (0, 0), (2, 2))
    t2 = sc.add_bitmap_texture(np.random.default_rng(9).random((8, 8, 3)).astype(np.float32))
    t3 = sc.add_grid_texture((0.9, 0.2, 0.2), (0.3, 0.3, 0.3), 0.05, (0, 0), (4, 1))
    sc.set_material(0, material_diffuse(kd_texture=t1))
    sc.set_material(2, material_diffuse(kd_texture=t2))
    m = material_substrate((0, 0, 0), (0.05, 0.05, 0.05), "ggx", 0.2)
    m.kd_texture = t3
    sc.set_material(5, m)
    dev, osc = DeviceScene(gpu_ctx, sc), ob.OracleScene(sc)
    for integ in (_abi.path_desc(), _abi.direct_desc(1, 1)):
        img, st = dev.render(integ, 6, seed=4)
        ref, so = osc.render(integ, 6, seed=4, cfg=ob.config(**STREAM))
        assert st.segments == so.segments and np.array_equal(img, ref)
    dev.close()


def test_constant_environment_bit_exact(gpu_ctx):
    """EnvironmentLight (constant): escaping rays add the environment with MIS, light sampling draws uniform directions and
    ends its shadow segment where the reference's BoundingSphere::intersect says (quirk kept), `direct` handles both misses."""
    sc = load_cbox(96, 96)
    sc.set_environment((0.3, 0.3, 0.4))
    sc.add_point_light((0.2, 0.2, 0.2), (0.0, 1.0, 0.5))
    dev, osc = DeviceScene(gpu_ctx, sc), ob.OracleScene(sc)
    for integ in (_abi.path_desc(), _abi.path_desc(strategy=_abi.RL_STRATEGY_EMITTER, max_depth=4), _abi.direct_desc(1, 1), _abi.direct_desc(2, 0)):
        img, st = dev.render(integ, 6, seed=11)
        ref, so = osc.render(integ, 6, seed=11, cfg=ob.config(**STREAM))
        assert (st.segments, st.hits, st.shadow_rays) == (so.segments, so.hits, so.shadow_rays)
        assert np.array_equal(img, ref)
    dev.close()


def test_config_shapes_c3_c5(gpu_ctx):
    """BASELINE configs[2] (Phong walls, 512x512) and configs[4] (1920x1080, Fov::Y quirk, ragged 16x16 tiles,
    material sort on): sub-sampled spp, bit-exact against the oracle on the same stream."""
    sc = load_cbox()
    kds = [(0.63, 0.065, 0.05), (0.14, 0.45, 0.091), (0.725, 0.71, 0.68)]
    for mesh, kd in [(0, kds[2]), (1, kds[2]), (2, kds[2]), (3, kds[1]), (4, kds[0])]:
        sc.set_material(mesh, material_phong([0.5 * c for c in kd], (0.3, 0.3, 0.3), 50.0))
    dev = DeviceScene(gpu_ctx, sc)
    img, st = dev.render(_abi.path_desc(), 2, seed=0)
    ref, so = ob.OracleScene(sc).render(_abi.path_desc(), 2, seed=0, cfg=ob.config(**STREAM))
    assert np.array_equal(img, ref) and st.segments == so.segments
    dev.close()
    wide = load_cbox(1920, 1080)
    dev = DeviceScene(gpu_ctx, wide)
    img, st = dev.render(_abi.path_desc(), 1, seed=0, material_sort=1)
    ref, so = ob.OracleScene(wide).render(_abi.path_desc(), 1, seed=0, cfg=ob.config(**STREAM))
    assert img.shape == (1080, 1920, 3) and np.array_equal(img, ref) and st.segments == so.segments
    dev.close()


def _with_tail(value, fn):
    """Run fn() with the k_tail hand-over threshold set (RL_TAIL_MAX is read on every rl_render; "0" = pure wavefront)."""
    old = os.environ.get("RL_TAIL_MAX")
    if value is None:
        os.environ.pop("RL_TAIL_MAX", None)
    else:
        os.environ["RL_TAIL_MAX"] = str(value)
    try:
        return fn()
    finally:
        if old is None:
            os.environ.pop("RL_TAIL_MAX", None)
        else:
            os.environ["RL_TAIL_MAX"] = old


@pytest.mark.parametrize("scene", ["cbox", "tess2", "mixed", "env"])
def test_tail_kernel_changes_nothing_but_the_schedule(gpu_ctx, scene):
    """k_tail (one thread follows one path to its end once the queue is short) against the pure wavefront: same image bits,
    same counters, fewer launches -- on the group table, the LBVH, every BSDF kind and the environment's miss edges.  The
    default threshold and "hand over as early as possible" both go through it; the oracle pins the result."""
    if scene == "cbox":
        sc = load_cbox(200, 136)
    elif scene == "tess2":
        import sys
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
        from tess_cbox import tessellated_cbox_json
        sc = SceneLoaderManager().load_string(tessellated_cbox_json(2), "json")
        sc.set_resolution(96, 96)
    elif scene == "mixed":
        from conftest import mixed_cbox
        sc = mixed_cbox(96, 96)
    else:
        sc = load_cbox(96, 96)
        sc.set_environment((0.4, 0.5, 0.7))
    dev = DeviceScene(gpu_ctx, sc)
    integ = _abi.path_desc()
    key = lambda st: (st.samples, st.segments, st.hits, st.shadow_rays, st.shadow_visible, st.max_depth_seen)
    a, sa = _with_tail(0, lambda: dev.render(integ, 12, seed=3, batch_spp=8))
    b, sb = _with_tail(None, lambda: dev.render(integ, 12, seed=3, batch_spp=8))
    c, sc_ = _with_tail(1 << 30, lambda: dev.render(integ, 12, seed=3, batch_spp=8))
    assert np.array_equal(a, b) and np.array_equal(a, c)
    assert key(sa) == key(sb) == key(sc_)
    assert sc_.kernel_launches <= sb.kernel_launches < sa.kernel_launches
    ref, so = ob.OracleScene(sc).render(integ, 12, seed=3, cfg=ob.config(**STREAM))
    assert np.array_equal(a, ref) and sa.segments == so.segments
    dev.close()


def test_endless_paths_are_an_error_not_a_hang(gpu_ctx):
    """Closed cube with albedo 1 and no Russian roulette: paths never end.  The wavefront loop gives up after 4095 iterations
    with RL_ERR_UNSUPPORTED; k_tail keeps the same limit and the same answer (it must not spin to the 16-bit depth guard)."""
    import json
    c = [(-1, -1, -1), (1, -1, -1), (1, 1, -1), (-1, 1, -1), (-1, -1, 1), (1, -1, 1), (1, 1, 1), (-1, 1, 1)]
    quads = [((0, 1, 2, 3), (0, 0, 1)), ((5, 4, 7, 6), (0, 0, -1)), ((4, 0, 3, 7), (1, 0, 0)), ((1, 5, 6, 2), (-1, 0, 0)),
             ((4, 5, 1, 0), (0, 1, 0)), ((3, 2, 6, 7), (0, -1, 0))]
    faces = [{"material": {"type": "diffuse", "kd": [1.0] * 3}, "emission": [1.0] * 3, "indices": [0, 1, 2, 0, 2, 3],
              "P": [x for i in q for x in c[i]], "N": list(n) * 4} for q, n in quads]
    sc = SceneLoaderManager().load_string(json.dumps({
        "camera": {"width": 8, "height": 8, "fov": 60, "to_world": [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, -1, 0, 0, 0, 0.5, 1]}, "meshes": faces}), "json")
    dev = DeviceScene(gpu_ctx, sc)
    integ = _abi.path_desc(rr_depth=1 << 20)  # roulette starts at depth 2^20: never (rr_depth=None would mean "always")
    for tail in (0, None):
        with pytest.raises(DeviceError) as e:
            _with_tail(tail, lambda: dev.render(integ, 1, seed=1))
        assert e.value.code == _abi.RL_ERR_UNSUPPORTED and "4095" in str(e.value)
    img, st = dev.render(_abi.path_desc(rr_depth=1 << 20, max_depth=40), 1, seed=1)  # bounded depth: fine
    assert st.max_depth_seen == 39 and np.isfinite(img).all()
    dev.close()


def test_pinned_host_buffer(cbox_dev):
    """rl_host_alloc / rl_host_free: rl_render into a page-locked buffer gives the same frame as into pageable memory."""
    from rustlight_b200.device import PinnedImage
    pin = PinnedImage(cbox_dev.height, cbox_dev.width)
    a, _ = cbox_dev.render(_abi.path_desc(), 2, seed=4)
    b, _ = cbox_dev.render(_abi.path_desc(), 2, seed=4, out=pin.array)
    assert b is pin.array and np.array_equal(a, b) and a.max() > 0
    b = None
    pin.close()
    lib().rl_host_free(None)


def test_consecutive_compute_calls_continue_the_sample_sequence(gpu_ctx):
    """Integrator::compute called twice on one sampler (what IntegratorAverage does, avg.rs:45-65) must not render the same
    samples twice: pass p covers samples [p * spp, (p + 1) * spp) -- two 8-spp passes average to the 16-spp image."""
    from rustlight_b200.device import IndependentSampler, IntegratorPathTracing
    sc = load_cbox(64, 64)
    dev = DeviceScene(gpu_ctx, sc)
    smp = IndependentSampler(5)
    integ = IntegratorPathTracing()
    a = integ.compute(smp, dev, spp=8).values["primal"]
    b = integ.compute(smp, dev, spp=8).values["primal"]
    assert smp.passes == 2 and not np.array_equal(a, b)
    full, _ = dev.render(integ.desc(), 16, seed=5)
    assert np.allclose(0.5 * (a + b), full, rtol=2e-6, atol=1e-7)
    dev.close()


def test_textures_on_every_colour_slot_bit_exact(gpu_ctx):
    """Ks / Kr / Kt / eta / k textures (bsdfs/mod.rs:218-253 applies bsdf_texture_match_pbrt to every colour parameter): the general
    shade kernel against the oracle, with and without the material sort."""
    from test_bsdfs import _cbox_textured_on_every_slot
    sc = _cbox_textured_on_every_slot(64, 64)
    dev, osc = DeviceScene(gpu_ctx, sc), ob.OracleScene(sc)
    for integ in (_abi.path_desc(), _abi.direct_desc(1, 1)):
        ref, so = osc.render(integ, 6, seed=9, cfg=ob.config(**STREAM))
        for sort in (0, 1):
            img, st = dev.render(integ, 6, seed=9, material_sort=sort)
            assert (st.segments, st.hits, st.shadow_rays) == (so.segments, so.hits, so.shadow_rays)
            assert np.array_equal(img, ref)
    dev.close()
