"""The light tree of `-x ats` (LightSamplerATS, emitter.rs:782-1400): Scene::build_emitters(true) replaces the flux-proportional
emitter choice by an importance-driven descent of a BVH over the emissive triangles.

(1) the oracle against the model: the branch probabilities form a distribution over the lights, sample() and pdf() agree, lights that
face away get no probability; (2) the device arithmetic (tests/emu: host-built tree of rl_ats_host.hpp + rl_device.cuh: ats_*) == the
oracle's independently typed restatement, bit for bit, on the tree functions and on renders; (3) statistics: with and without the tree
the image converges to the same values, with less noise where lights are far apart.  GPU: test_gpu.py."""
import json
import os
import sys

import numpy as np
import pytest

import emu_binding as eb
from conftest import ROOT, load_cbox, rel_l2
from oracle import binding as ob
from rustlight_b200 import SceneLoaderManager, _abi
from rustlight_b200.host import SceneError

sys.path.insert(0, os.path.join(ROOT, "tools"))
STREAM = dict(estimator=ob.EST_STREAM, accel_mode=ob.ACCEL_BVH)


def many_lights_scene(w=32, h=32, ats=True, seed=0, nlights=9):
    """The Cornell box with its ceiling light tessellated (8 triangles) plus small emissive quads of different colours, sizes and
    orientations scattered in the box: ~26 emissive triangles in 10 meshes."""
    from tess_cbox import tessellated_cbox_json
    base = json.loads(tessellated_cbox_json(2))
    rng = np.random.default_rng(seed)
    for k in range(nlights):
        c = rng.uniform([-0.8, 0.2, -0.8], [0.8, 1.8, 0.8])
        a, b = rng.normal(size=3), rng.normal(size=3)
        a *= rng.uniform(0.03, 0.15) / np.linalg.norm(a)
        b -= a * (a @ b) / (a @ a)
        b *= rng.uniform(0.03, 0.15) / np.linalg.norm(b)
        P = np.array([c - a - b, c + a - b, c + a + b, c - a + b], np.float32)
        # vertex normals along (v1 - v0) x (v2 - v0): the axis convert_light_proxy gives the orientation cone (emitter.rs:739, 753).  Without
        # them sample_tri emits towards (v2 - v0) x (v1 - v0) (geometry.rs:272-276), the OPPOSITE side, and the tree assigns such a lamp zero
        # importance exactly where it shines -- a quirk of the reference that is reproduced, but not one a convergence test can use.
        nrm = np.cross(P[1] - P[0], P[2] - P[0])
        nrm /= np.linalg.norm(nrm)
        base["meshes"].append({"name": f"lamp{k}", "material": {"type": "diffuse", "kd": [0, 0, 0]}, "emission": [float(x) for x in rng.uniform(2, 30, 3)],
                               "indices": [0, 1, 2, 0, 2, 3], "P": [float(x) for x in P.ravel()], "N": [float(x) for x in np.tile(nrm, 4)]})
    base["camera"]["width"], base["camera"]["height"] = w, h
    sc = SceneLoaderManager().load_string(json.dumps(base), "json")
    sc.set_resolution(w, h)
    return sc.set_ats(ats)


def _light_prims(sc):
    d = sc.desc.contents
    out, first = [], 0
    for i in range(d.nmeshes):
        if d.meshes[i].emission_kind:
            out += list(range(first, first + d.meshes[i].ntris))
        first += d.meshes[i].ntris
    return out


def _points(n, seed):
    rng = np.random.default_rng(seed)
    p = rng.uniform([-0.95, 0.05, -0.95], [0.95, 1.9, 0.95], (n, 3)).astype(np.float32)
    nn = rng.normal(size=(n, 3))
    return p, (nn / np.linalg.norm(nn, axis=1, keepdims=True)).astype(np.float32)


# ---- (1) the oracle against the model ------------------------------------------------------------------------------------------
def test_tree_probabilities_form_a_distribution():
    sc = many_lights_scene()
    osc = ob.OracleScene(sc)
    prims = _light_prims(sc)
    assert len(prims) == 8 + 2 * 9
    P, N = _points(40, 1)
    for p, n in zip(P, N):
        for nn in (None, n):
            tot = sum(osc.ats_pdf(q, p, nn) for q in prims)
            assert tot == pytest.approx(1.0, abs=2e-5)
            for r in np.random.default_rng(2).random(12).astype(np.float32):
                prim, pdf = osc.ats_sample(r, p, nn)
                assert prim in prims and pdf > 0 and pdf == pytest.approx(osc.ats_pdf(prim, p, nn), rel=1e-5)


def test_sampling_frequencies_follow_the_pdf():
    sc = many_lights_scene()
    osc = ob.OracleScene(sc)
    prims = _light_prims(sc)
    p, n = np.float32([0.1, 0.4, 0.2]), np.float32([0, 1, 0])
    pdf = np.array([osc.ats_pdf(q, p, n) for q in prims])
    rs = (np.arange(20000) + 0.5) / 20000
    counts = np.zeros(len(prims))
    for r in rs.astype(np.float32):
        counts[prims.index(osc.ats_sample(r, p, n)[0])] += 1
    assert np.allclose(counts / len(rs), pdf, atol=2e-3)
    assert (pdf > 0).sum() >= 6 and (pdf == 0).sum() >= 1  # lamps that face away from p get no probability (theta_e = pi / 2)


def test_only_mesh_emitters_and_direct_restrictions():
    sc = many_lights_scene(16, 16)
    sc.add_point_light((1, 1, 1), (0, 1, 0))
    with pytest.raises(RuntimeError, match="surface"):
        ob.OracleScene(sc)
    with pytest.raises(RuntimeError, match="surface"):
        eb.EmuScene(sc)
    dark = load_cbox(16, 16)
    dark.set_ats(True)
    assert ob.OracleScene(dark) is not None  # the Cornell box itself: two emissive triangles


# ---- (2) device arithmetic == oracle, bit for bit -----------------------------------------------------------------------------------
def test_device_tree_functions_bit_exact():
    sc = many_lights_scene(seed=3)
    esc, osc = eb.EmuScene(sc), ob.OracleScene(sc)
    prims = _light_prims(sc)
    P, N = _points(300, 4)
    rs = np.random.default_rng(5).random(300).astype(np.float32)
    for p, n, r in zip(P, N, rs):
        assert esc.ats_sample(r, p, n) == osc.ats_sample(r, p, n)
        q = prims[int(r * len(prims))]
        assert esc.ats_pdf(q, p, n) == osc.ats_pdf(q, p, n) and esc.ats_pdf(q, p, None) == osc.ats_pdf(q, p, None)


@pytest.mark.parametrize("integ", [_abi.path_desc(), _abi.path_desc(strategy=_abi.RL_STRATEGY_EMITTER), _abi.path_desc(max_depth=4, rr_depth=2),
                                   _abi.direct_desc(0, 2), _abi.direct_desc(1, 1), _abi.direct_desc(2, 3)],
                         ids=["path", "path-emitter", "path-d4", "direct02", "direct11", "direct23"])
def test_light_tree_render_bit_exact(integ):
    sc = many_lights_scene(32, 32)
    ie, se = eb.EmuScene(sc, "sah4").render(integ, 4, seed=4)
    io, so = ob.OracleScene(sc).render(integ, 4, seed=4, cfg=ob.config(**STREAM))
    assert np.isfinite(io).all() and io.mean() > 0.02
    assert (se.segments, se.hits, se.shadow_rays) == (so.segments, so.hits, so.shadow_rays)
    assert np.array_equal(ie, io)


def test_light_tree_stream_estimator_equals_graph():
    """The MIS of path.rs with the tree: light samples use the tree with Some(n_s), the pdf of a BSDF-sampled hit on a light uses it
    with None (emitters.rs:52-57) -- both estimators must take the same decisions."""
    osc = ob.OracleScene(many_lights_scene(24, 24))
    integ = _abi.path_desc(max_depth=5)
    a, sa = osc.render(integ, 8, seed=2, cfg=ob.config(**STREAM))
    b, sb = osc.render(integ, 8, seed=2, cfg=ob.config(estimator=ob.EST_GRAPH, accel_mode=ob.ACCEL_BVH))
    assert (sa.segments, sa.shadow_rays) == (sb.segments, sb.shadow_rays) and rel_l2(a, b) < 1e-6


# ---- (3) statistics -----------------------------------------------------------------------------------------------------------------------
def test_tree_and_flux_sampling_converge_to_the_same_image():
    integ = _abi.direct_desc(0, 1)
    a, _ = ob.OracleScene(many_lights_scene(16, 16, ats=True)).render(integ, 400, seed=1, cfg=ob.config(**STREAM))
    b, _ = ob.OracleScene(many_lights_scene(16, 16, ats=False)).render(integ, 400, seed=1, cfg=ob.config(**STREAM))
    assert abs(a.mean() - b.mean()) < 0.03 * b.mean()
    ref, _ = ob.OracleScene(many_lights_scene(16, 16, ats=False)).render(integ, 3000, seed=9, cfg=ob.config(**STREAM))
    assert rel_l2(a, ref) < rel_l2(b, ref)  # the tree places samples where the light comes from: less noise at equal spp
