"""SURVEY 8(d) CPU baseline for every BASELINE.json config: the oracle in its reference-faithful configuration (graph estimator,
BVHAccel, glibc math, per-block xoshiro streams = sampler mode A) on the host cores of the box, all logical cores and one
thread.  C1 runs exactly; C2-C5 run a reduced sample count and are scaled linearly in spp (cost per sample does not depend
on spp: every pixel sample is an independent path).  Timed region = block loop + merge ("Elapsed Integrator",
integrators/mod.rs:323-334).  Test infrastructure: nothing of the product runs here."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_cbox  # noqa: E402
from oracle import binding as ob  # noqa: E402
from rustlight_b200 import _abi  # noqa: E402
from rustlight_b200.host import material_phong  # noqa: E402


def run(name, sc, integ, spp_full, spp_run, threads):
    osc = ob.OracleScene(sc)
    cfg = ob.config(math_mode=ob.MATH_LIBM, accel_mode=ob.ACCEL_BVH, estimator=ob.EST_GRAPH, nthreads=threads)
    _, st = osc.render(integ, spp_run, seed=0, sampler_mode=_abi.RL_SAMPLER_BLOCK_STREAM, cfg=cfg)
    full_s = st.seconds * spp_full / spp_run
    print(json.dumps({"config": name, "threads": int(st.threads_used), "spp_run": spp_run, "spp_full": spp_full, "seconds_run": st.seconds,
                      "seconds_full" + ("" if spp_run == spp_full else "_extrapolated"): full_s,
                      "Msamples/s": st.samples / st.seconds / 1e6, "Msegments/s": st.segments / st.seconds / 1e6}), flush=True)


phong = load_cbox(512, 512)
kds = [(0.63, 0.065, 0.05), (0.14, 0.45, 0.091), (0.725, 0.71, 0.68)]
for mesh, kd in [(0, kds[2]), (1, kds[2]), (2, kds[2]), (3, kds[1]), (4, kds[0])]:
    phong.set_material(mesh, material_phong([0.5 * c for c in kd], (0.3, 0.3, 0.3), 50.0))
for threads in (0, 1):  # 0 = all logical cores
    few = 1 if threads == 1 else 4
    run("C1 cbox path 512x512 x16", load_cbox(512, 512), _abi.path_desc(), 16, 16, threads)
    run("C2 cbox path 1024x1024 x128", load_cbox(1024, 1024), _abi.path_desc(), 128, few * 2, threads)
    run("C3 Phong walls path 512x512 x512", phong, _abi.path_desc(), 512, few * 4, threads)
    run("C4 cbox direct -b 1 -l 1 2048x2048 x64", load_cbox(2048, 2048), _abi.direct_desc(1, 1), 64, few, threads)
    run("C5 cbox path 1920x1080 x4096", load_cbox(1920, 1080), _abi.path_desc(), 4096, few, threads)
