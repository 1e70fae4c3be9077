"""Generates tests/golden/*.npz from the oracle (run in the build container).

The reference ships no golden vectors and cannot run here (F2), so these
fixtures are outputs of the oracle in its reference-faithful configuration (path GRAPH
estimator, BVHAccel, glibc math) for sampler mode B, plus hit indices of the fixed
primary-ray grid.  They pin the oracle against regressions and give the GPU tests a target
that does not need the oracle binary.
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import load_cbox  # noqa: E402
from oracle import binding as ob  # noqa: E402
from rustlight_b200 import _abi  # noqa: E402


def main():
    faithful = ob.config(math_mode=ob.MATH_LIBM, accel_mode=ob.ACCEL_BVH, estimator=ob.EST_GRAPH)
    sc = load_cbox(64, 64)
    osc = ob.OracleScene(sc)
    out = {}
    for name, integ in [("path", _abi.path_desc()), ("path_bsdf", _abi.path_desc(strategy=_abi.RL_STRATEGY_BSDF)),
                        ("path_d4", _abi.path_desc(max_depth=4)), ("direct", _abi.direct_desc(1, 1))]:
        img, st = osc.render(integ, 16, seed=0, cfg=faithful)
        out[name] = img
        out[name + "_counts"] = np.array([st.segments, st.shadow_rays, st.hits], np.uint64)
    np.savez_compressed(os.path.join(HERE, "cbox64_spp16_seed0.npz"), **out)
    full = ob.OracleScene(load_cbox())
    prim, tuv = full.primary_hits(ob.ACCEL_BVH)  # the reference's default accelerator (visit order decides exact ties)
    prim8 = np.where(prim == 0xFFFFFFFF, 255, prim).astype(np.uint8)  # 36 triangles; 255 = miss
    np.savez_compressed(os.path.join(HERE, "cbox512_primary_hits.npz"), prim=prim8, t_sub8=tuv[::8, ::8, 0])
    print({k: (v.shape, float(v.mean())) for k, v in out.items()})


def emission():
    """EmissionType::HSV / Texture lamps (-x hvs-light / -x texture-light), same configuration: cbox48_emission_spp8_seed0.npz."""
    from test_emission import hsv_box, texture_box
    faithful = ob.config(math_mode=ob.MATH_LIBM, accel_mode=ob.ACCEL_BVH, estimator=ob.EST_GRAPH)
    out = {}
    for name, make, integ in [("hsv_path", hsv_box, _abi.path_desc()), ("hsv_direct", hsv_box, _abi.direct_desc(1, 1)),
                              ("texture_path", texture_box, _abi.path_desc()), ("texture_direct", texture_box, _abi.direct_desc(1, 1))]:
        img, st = ob.OracleScene(make(48, 48)).render(integ, 8, seed=0, cfg=faithful)
        out[name] = img
        out[name + "_counts"] = np.array([st.segments, st.shadow_rays, st.hits], np.uint64)
    np.savez_compressed(os.path.join(HERE, "cbox48_emission_spp8_seed0.npz"), **out)
    print({k: (v.shape, float(v.mean())) for k, v in out.items()})


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "emission":
        emission()  # (added in round 3; the older files are left byte for byte as they were)
    else:
        main()
