"""Quick on-GPU parity + timing check (development aid; the real tests are tests/ -m gpu)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import binding as ob  # noqa: E402
from rustlight_b200 import SceneLoaderManager, _abi  # noqa: E402
from rustlight_b200.device import Context, DeviceScene  # noqa: E402


def main():
    scene = SceneLoaderManager().load(os.path.join(ROOT, "data", "cbox.pbrt"))
    ctx = Context(0)
    t0 = time.time()
    dsc = DeviceScene(ctx, scene)
    bi = dsc.bvh_info()
    print("scene create %.1f ms; bvh: ntris=%d nnodes=%d depth=%d smem=%d" % ((time.time() - t0) * 1e3, bi.ntris, bi.nnodes, bi.max_depth, bi.smem_resident))
    osc = ob.OracleScene(scene)
    pg, tg = dsc.primary_hits()
    po, to = osc.primary_hits(ob.ACCEL_BVH)
    print("primary hits: prim mismatches", int((pg != po).sum()), "tuv bit-exact", bool(np.array_equal(tg, to)))
    rng = np.random.default_rng(0)
    n = 200000
    o = rng.uniform(-0.99, 0.99, (n, 3)).astype(np.float32) + np.array([0, 1, 0], np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d = d.astype(np.float32)
    pg, tg = dsc.trace(o, d)
    po, to = osc.trace(o, d, ob.ACCEL_BVH)
    print("random rays: prim mismatches", int((pg != po).sum()), "tuv bit-exact", bool(np.array_equal(tg, to)))
    p1 = rng.uniform(-0.99, 0.99, (n, 3)).astype(np.float32) + np.array([0, 1, 0], np.float32)
    vg = dsc.visible(o, p1)
    vo = osc.visible(o, p1, ob.ACCEL_BVH)
    print("visible: mismatches", int((vg != vo).sum()), "visible frac", float(vg.mean()))
    integ = _abi.path_desc()
    img, st = dsc.render(integ, 4, seed=0)
    ref, ost = osc.render(integ, 4, seed=0, cfg=ob.config(estimator=ob.EST_STREAM, accel_mode=ob.ACCEL_BVH))
    print("render 512x512x4: gpu", {k: v for k, v in st.as_dict().items() if not k.startswith("ms_") or v})
    print("                  orc", ost.as_dict())
    print("  bit-exact:", bool(np.array_equal(img, ref)), "ndiff px", int((img != ref).any(axis=2).sum()), "rel_l2", float(np.linalg.norm(img - ref) / np.linalg.norm(ref)))
    # timing
    for spp, (w, h) in [(16, (512, 512)), (128, (1024, 1024))]:
        scene.set_resolution(w, h)
        d2 = DeviceScene(ctx, scene)
        d2.render(integ, 2, seed=0, want_image=False)
        _, st = d2.render(integ, spp, seed=0, want_image=False)
        print(json.dumps({"w": w, "h": h, "spp": spp, "ms": st.ms_total, "Msamples/s": st.samples / st.ms_total / 1e3,
                          "Mseg/s": st.segments / st.ms_total / 1e3, "launches": st.kernel_launches, "max_depth": st.max_depth_seen}))
        ctx.set_profiling(True)
        _, st = d2.render(integ, spp, seed=0, want_image=False)
        ctx.set_profiling(False)
        print("  profile ms:", {k: round(v, 2) for k, v in st.as_dict().items() if k.startswith("ms_")})
        d2.close()


if __name__ == "__main__":
    main()
