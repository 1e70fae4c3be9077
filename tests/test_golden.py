"""Oracle vs the committed golden fixtures (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_cbox, rel_l2
from oracle import binding as ob
from rustlight_b200 import _abi

INTEGS = {"path": _abi.path_desc(), "path_bsdf": _abi.path_desc(strategy=_abi.RL_STRATEGY_BSDF),
          "path_d4": _abi.path_desc(max_depth=4), "direct": _abi.direct_desc(1, 1)}


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(GOLDEN, "cbox64_spp16_seed0.npz"))


@pytest.mark.parametrize("name", list(INTEGS))
def test_oracle_reproduces_golden(golden, name):
    osc = ob.OracleScene(load_cbox(64, 64))
    img, st = osc.render(INTEGS[name], 16, seed=0, cfg=ob.config(math_mode=ob.MATH_LIBM, accel_mode=ob.ACCEL_BVH, estimator=ob.EST_GRAPH))
    assert np.array_equal(img, golden[name])
    assert [st.segments, st.shadow_rays, st.hits] == list(golden[name + "_counts"])


@pytest.mark.parametrize("name", ["path", "path_bsdf", "path_d4"])
def test_gpu_configuration_of_the_oracle_is_within_tolerance_of_golden(golden, name):
    """spec math + brute-force tie order + stream estimator (what the GPU computes) vs the faithful one."""
    osc = ob.OracleScene(load_cbox(64, 64))
    img, st = osc.render(INTEGS[name], 16, seed=0, cfg=ob.config(math_mode=ob.MATH_SPEC, accel_mode=ob.ACCEL_BVH, estimator=ob.EST_STREAM))
    assert rel_l2(img, golden[name]) < 1e-3      # north_star tolerance; typically ~1e-6
    assert abs(int(st.segments) - int(golden[name + "_counts"][0])) <= 4


def test_primary_hits_golden():
    g = np.load(os.path.join(GOLDEN, "cbox512_primary_hits.npz"))
    prim, tuv = ob.OracleScene(load_cbox()).primary_hits(ob.ACCEL_BVH)
    assert np.array_equal(np.where(prim == 0xFFFFFFFF, 255, prim).astype(np.uint8), g["prim"])
    assert np.array_equal(tuv[::8, ::8, 0], g["t_sub8"])


@pytest.mark.parametrize("name", ["hsv_path", "hsv_direct", "texture_path", "texture_direct"])
def test_oracle_reproduces_emission_golden(name):
    """uv-dependent emission (EmissionType::HSV / Texture): tests/golden/make_golden.py emission."""
    from test_emission import hsv_box, texture_box
    g = np.load(os.path.join(GOLDEN, "cbox48_emission_spp8_seed0.npz"))
    make = hsv_box if name.startswith("hsv") else texture_box
    integ = _abi.path_desc() if name.endswith("path") else _abi.direct_desc(1, 1)
    img, st = ob.OracleScene(make(48, 48)).render(integ, 8, seed=0, cfg=ob.config(math_mode=ob.MATH_LIBM, accel_mode=ob.ACCEL_BVH, estimator=ob.EST_GRAPH))
    assert np.array_equal(img, g[name])
    assert [st.segments, st.shadow_rays, st.hits] == list(g[name + "_counts"])
