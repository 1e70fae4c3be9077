"""Device arithmetic (rl_device.cuh / rl_build.cuh compiled for the CPU by tests/emu) vs the oracle.

This is the CPU-side half of the parity argument: the functions the kernels call take the same
discrete decisions, and produce the same radiance bits, as the oracle's stream estimator.  The
GPU half (tests/test_gpu.py) checks that the kernels around them move the records correctly.
"""
import os

import numpy as np
import pytest

import emu_binding as eb
from conftest import load_cbox, soup_scene
from oracle import binding as ob
from rustlight_b200 import SceneLoaderManager, _abi
from rustlight_b200.host import material_phong

STREAM = dict(estimator=ob.EST_STREAM, accel_mode=ob.ACCEL_BVH)


def _rays(n, seed, lo=-0.99, hi=0.99, shift=(0, 1, 0)):
    rng = np.random.default_rng(seed)
    o = (rng.uniform(lo, hi, (n, 3)) + shift).astype(np.float32)
    d = rng.normal(size=(n, 3))
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    p1 = (rng.uniform(lo, hi, (n, 3)) + shift).astype(np.float32)
    return o, d, p1


def test_lbvh_is_valid(cbox):
    esc = eb.EmuScene(cbox)
    assert esc.bvh_validate() == 0 and 1 <= esc.bvh_max_depth() <= 36


def test_cbox_group_table_pairs_every_quad(cbox):
    """36 triangles = 18 planar quads -> 18 pair records in 9 groups; the plane mismatch inside a pair is rounding noise."""
    info = eb.EmuScene(cbox).flat_info()
    assert info["groups"] == 9 and info["pairs"] == 18 and info["singles"] == 0 and info["delta"] < 1e-6


@pytest.mark.parametrize("accel", ["flat", "leaf", "tree", "sah2", "sah4"])
def test_accel_modes_agree_with_the_oracle(cbox, cbox_oracle, accel):
    """The traversal modes of the device (group table, one big leaf, Morton LBVH, 2- and 4-wide tree over the reference's topology)
    give the oracle's hits bit for bit."""
    esc = eb.EmuScene(cbox, accel)
    assert (esc.flat_info()["groups"] > 0) == (accel == "flat")
    o, d, p1 = _rays(20000, 3)
    pe, te = esc.trace(o, d)
    po, to = cbox_oracle.trace(o, d, ob.ACCEL_BVH)
    assert np.array_equal(pe, po) and np.array_equal(te, to)
    assert np.array_equal(esc.visible(o, p1), cbox_oracle.visible(o, p1, ob.ACCEL_BVH))
    # rays that start ON the surfaces (what the path tracer actually traces), incl. toward the light
    hit = po != 0xFFFFFFFF
    o2 = (o[hit] + d[hit] * to[hit, :1]).astype(np.float32)
    d2 = _rays(len(o2), 4)[1]
    pe, te = esc.trace(o2, d2)
    po2, to2 = cbox_oracle.trace(o2, d2, ob.ACCEL_BVH)
    assert np.array_equal(pe, po2) and np.array_equal(te, to2)
    light = np.tile(np.array([[0.0, 1.98, 0.0]], np.float32), (len(o2), 1)) + _rays(len(o2), 5)[0] * np.float32(0.2) * [1, 0, 1]
    light = light.astype(np.float32)
    assert np.array_equal(esc.visible(o2, light), cbox_oracle.visible(o2, light, ob.ACCEL_BVH))


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_pair_records_are_conservative_on_adversarial_quads(seed):
    from conftest import adversarial_pairs_case
    sc, o, dd, p1 = adversarial_pairs_case(seed)
    esc, osc = eb.EmuScene(sc), ob.OracleScene(sc)
    info = esc.flat_info()
    assert info["groups"] > 0 and info["pairs"] >= 3, info
    pe, te = esc.trace(o, dd)
    po, to = osc.trace(o, dd, ob.ACCEL_BVH)
    assert (po != 0xFFFFFFFF).mean() > 0.3
    assert np.array_equal(pe, po) and np.array_equal(te, to)
    assert np.array_equal(esc.visible(o, p1), osc.visible(o, p1, ob.ACCEL_BVH))


def test_primary_hits_exact(cbox, cbox_oracle):
    pe, te = eb.EmuScene(cbox).primary_hits()
    po, to = cbox_oracle.primary_hits(ob.ACCEL_BVH)
    assert np.array_equal(pe, po) and np.array_equal(te, to)


def test_random_rays_and_segments_exact(cbox, cbox_oracle):
    esc = eb.EmuScene(cbox)
    o, d, p1 = _rays(30000, 1)
    pe, te = esc.trace(o, d)
    po, to = cbox_oracle.trace(o, d, ob.ACCEL_BVH)
    assert np.array_equal(pe, po) and np.array_equal(te, to)
    assert np.array_equal(esc.visible(o, p1), cbox_oracle.visible(o, p1, ob.ACCEL_BVH))


@pytest.mark.parametrize("accel", ["tree", "sah2", "sah4"])
@pytest.mark.parametrize("ntris,seed", [(50, 0), (600, 1), (3000, 2)])
def test_soup_trees_equal_the_oracle(ntris, seed, accel):
    """Conservative culling: neither the Morton LBVH nor the 2- / 4-wide trees built over the reference's topology (rl_wide_host.hpp:
    what the device walks on scenes without a group table) lose a triangle the exact test accepts; every triangle sits in exactly
    one leaf and every child box contains its subtree."""
    sc = SceneLoaderManager().load_string(soup_scene(ntris, seed), "json")
    esc, osc = eb.EmuScene(sc, accel), ob.OracleScene(sc)
    assert esc.bvh_validate() == 0 and esc.bvh_max_depth() >= 2
    o, d, p1 = _rays(6000, seed + 10, -1.2, 1.2, (0, 0, 0))
    pe, te = esc.trace(o, d)
    po, to = osc.trace(o, d, ob.ACCEL_BVH)
    assert np.array_equal(pe, po) and np.array_equal(te, to)
    assert np.array_equal(esc.visible(o, p1), osc.visible(o, p1, ob.ACCEL_BVH))


def test_far_origin_rays_exact(cbox, cbox_oracle):
    # origins far from the scene stress the culling epsilon (scaled by |o|)
    esc = eb.EmuScene(cbox)
    rng = np.random.default_rng(7)
    tgt = (rng.uniform(-1, 1, (4000, 3)) + [0, 1, 0])
    o = (tgt + rng.normal(size=(4000, 3)) * 300).astype(np.float32)
    d = tgt - o
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    pe, te = esc.trace(o, d)
    po, to = cbox_oracle.trace(o, d, ob.ACCEL_BVH)
    assert np.array_equal(pe, po) and np.array_equal(te, to)


def test_axis_aligned_rays_exact(cbox, cbox_oracle):
    # zero direction components: 1/0 = inf in the slab tests
    esc = eb.EmuScene(cbox)
    rng = np.random.default_rng(8)
    o = (rng.uniform(-0.9, 0.9, (3000, 3)) + [0, 1, 0]).astype(np.float32)
    axes = np.eye(3, dtype=np.float32)[rng.integers(0, 3, 3000)] * rng.choice([-1, 1], (3000, 1)).astype(np.float32)
    pe, te = esc.trace(o, axes)
    po, to = cbox_oracle.trace(o, axes, ob.ACCEL_BVH)
    assert np.array_equal(pe, po) and np.array_equal(te, to)


@pytest.mark.parametrize("kw", [dict(), dict(strategy=_abi.RL_STRATEGY_BSDF), dict(strategy=_abi.RL_STRATEGY_EMITTER),
                                dict(max_depth=2), dict(max_depth=4, min_depth=2), dict(rr_depth=4, max_depth=9),
                                dict(rr_depth=None), dict(single_scattering=True)])
def test_path_render_bit_exact(kw):
    sc = load_cbox(96, 96)
    integ = _abi.path_desc(**kw)
    ie, se = eb.EmuScene(sc).render(integ, 6, seed=5)
    io, so = ob.OracleScene(sc).render(integ, 6, seed=5, cfg=ob.config(**STREAM))
    assert (se.segments, se.hits, se.shadow_rays, se.shadow_visible) == (so.segments, so.hits, so.shadow_rays, so.nee_added)
    assert np.array_equal(ie, io)


def test_phong_render_bit_exact():
    sc = load_cbox(64, 64)
    kds = [(0.63, 0.065, 0.05), (0.14, 0.45, 0.091), (0.725, 0.71, 0.68)]
    for mesh, kd in [(0, kds[2]), (1, kds[2]), (2, kds[2]), (3, kds[1]), (4, kds[0])]:
        sc.set_material(mesh, material_phong([0.5 * c for c in kd], (0.3, 0.3, 0.3), 50.0))
    integ = _abi.path_desc()
    ie, se = eb.EmuScene(sc).render(integ, 8, seed=2)
    io, so = ob.OracleScene(sc).render(integ, 8, seed=2, cfg=ob.config(**STREAM))
    assert (se.segments, se.hits, se.shadow_rays) == (so.segments, so.hits, so.shadow_rays)
    assert np.array_equal(ie, io)


def test_soup_render_bit_exact():
    sc = SceneLoaderManager().load_string(soup_scene(400, 4, 48, 48), "json")
    integ = _abi.path_desc(max_depth=6)
    ie, se = eb.EmuScene(sc).render(integ, 4, seed=1)
    io, so = ob.OracleScene(sc).render(integ, 4, seed=1, cfg=ob.config(**STREAM))
    assert (se.segments, se.hits, se.shadow_rays) == (so.segments, so.hits, so.shadow_rays)
    assert np.array_equal(ie, io)


def test_tile_partition(cbox64):
    esc = eb.EmuScene(cbox64)
    integ = _abi.path_desc()
    full, _ = esc.render(integ, 3)
    parts = [esc.render(integ, 3, rank=r, nranks=2)[0] for r in range(2)]
    assert np.array_equal(parts[0] + parts[1], full)
    assert not (parts[0].astype(bool) & parts[1].astype(bool)).any()


@pytest.mark.parametrize("nb,nl", [(1, 1), (2, 3), (0, 2), (3, 0)])
def test_direct_render_bit_exact(nb, nl):
    sc = load_cbox(96, 80)
    integ = _abi.direct_desc(nb, nl)
    ie, se = eb.EmuScene(sc).render(integ, 5, seed=4)
    io, so = ob.OracleScene(sc).render(integ, 5, seed=4, cfg=ob.config(accel_mode=ob.ACCEL_BVH))
    assert (se.segments, se.hits, se.shadow_rays) == (so.segments, so.hits, so.shadow_rays)
    assert np.array_equal(ie, io)


@pytest.mark.parametrize("seed", range(6))
def test_prefilter_is_conservative_on_adversarial_scenes(seed):
    """Tiny, huge, sliver and far-away triangles; grazing, edge-on and vertex-on rays.  The leaf prefilter
    (fma, approximate reciprocal, margins) must never reject what the exact test accepts."""
    from conftest import adversarial_case
    sc, o, dd, p1 = adversarial_case(seed)
    esc, osc = eb.EmuScene(sc), ob.OracleScene(sc)
    pe, te = esc.trace(o, dd)
    po, to = osc.trace(o, dd, ob.ACCEL_BVH)
    assert np.array_equal(pe, po) and np.array_equal(te, to)
    assert (po != 0xFFFFFFFF).mean() > 0.2  # the rays do hit things
    assert np.array_equal(esc.visible(o, p1), osc.visible(o, p1, ob.ACCEL_BVH))


@pytest.mark.parametrize("dist,nc", [(1.0, False), (None, False), (0.3, True)])
def test_ao_render_bit_exact(dist, nc):
    """IntegratorAO (ao.rs): values are 0 or 1 per sample, so the images agree exactly iff every hit/miss and every
    distance comparison does."""
    sc = load_cbox(64, 64)
    integ = _abi.ao_desc(dist, nc)
    ie, se = eb.EmuScene(sc).render(integ, 8, seed=3)
    io, so = ob.OracleScene(sc).render(integ, 8, seed=3, cfg=ob.config(**STREAM))
    assert se.segments == so.segments and np.array_equal(ie, io)
    assert 0.05 < io.mean() < 0.95 and set(np.unique(io * 8).tolist()) <= set(range(9))


def test_prefilter_margins_carry_tenfold_slack(tmp_path):
    """The group-table prefilter may only reject what the exact test rejects.  Rebuilding the emulator with every margin
    scaled by 0.1 (RL_FLAT_MARGIN_SCALE) must still reproduce the oracle on the adversarial quads and on rays that start on
    the Cornell box surfaces: the shipped margins carry at least 10x slack (first mismatches appear at 0.01x)."""
    import ctypes
    import subprocess
    from conftest import ROOT, adversarial_pairs_case
    so = str(tmp_path / "libemu_tight.so")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-DRL_FLAT_MARGIN_SCALE=0.1f",
                           "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "rustlight_b200", "csrc"), "-shared", "-x", "c++",
                           os.path.join(ROOT, "tests", "emu", "emu.cpp"), "-o", so], stderr=subprocess.DEVNULL)
    L = ctypes.CDLL(so)
    L.emu_scene_create.restype = ctypes.c_void_p
    L.emu_scene_create.argtypes = [ctypes.POINTER(_abi.rl_scene_desc), ctypes.c_char_p, ctypes.c_size_t]
    L.emu_trace.argtypes = [ctypes.c_void_p, ctypes.c_size_t, eb.FP, eb.FP, ctypes.POINTER(ctypes.c_uint32), eb.FP]
    L.emu_scene_destroy.argtypes = [ctypes.c_void_p]

    def trace(sc, o, d):
        h = L.emu_scene_create(sc.desc, None, 0)
        prim, tuv = np.zeros(len(o), np.uint32), np.zeros((len(o), 3), np.float32)
        L.emu_trace(h, len(o), eb._f(o), eb._f(d), prim.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), eb._f(tuv))
        L.emu_scene_destroy(h)
        return prim, tuv
    for seed in range(4):
        sc, o, dd, _ = adversarial_pairs_case(seed)
        pe, te = trace(sc, o, dd)
        po, to = ob.OracleScene(sc).trace(o, dd, ob.ACCEL_BVH)
        assert np.array_equal(pe, po) and np.array_equal(te, to)
    sc = load_cbox()
    osc = ob.OracleScene(sc)
    o, d, _ = _rays(60000, 31)
    po, to = osc.trace(o, d, ob.ACCEL_BVH)
    hit = po != 0xFFFFFFFF
    o2 = np.ascontiguousarray((o[hit] + d[hit] * to[hit, :1]).astype(np.float32))
    d2 = np.ascontiguousarray(_rays(len(o2), 32)[1])
    pe, te = trace(sc, o2, d2)
    po2, to2 = osc.trace(o2, d2, ob.ACCEL_BVH)
    assert np.array_equal(pe, po2) and np.array_equal(te, to2)
