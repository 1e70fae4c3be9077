"""One rank's share of config 2 on one GPU (Context(0, nranks=N, rank=r) without a communicator): per-rank time, stage
times and load balance across ranks -- a cheap proxy for the N-GPU run (development aid).  usage: ab_rank.py N [spp]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rustlight_b200 import SceneLoaderManager, _abi  # noqa: E402
from rustlight_b200.device import Context, DeviceScene  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 128
sc = SceneLoaderManager().load(os.path.join(ROOT, "data", "cbox.pbrt")).scale_image(2.0)
integ = _abi.path_desc()
full = Context(0)
dev = DeviceScene(full, sc)
dev.render(integ, 8, want_image=False)
t_full = min(dev.render(integ, spp, want_image=False)[1].ms_total for _ in range(3))
full.close()
rows = []
for r in range(n):
    ctx = Context(0, nranks=n, rank=r)
    dev = DeviceScene(ctx, sc)
    dev.render(integ, 8, want_image=False)
    sts = [dev.render(integ, spp, want_image=False)[1] for _ in range(3)]
    best = min(sts, key=lambda s: s.ms_total)
    ctx.set_profiling(True)
    _, st = dev.render(integ, spp, want_image=False)
    rows.append({"rank": r, "ms": best.ms_total, "segments": best.segments, "launches": best.kernel_launches, "iters": best.max_depth_seen,
                 "trace": st.ms_trace, "shade": st.ms_shade, "shadow": st.ms_shadow})
    ctx.close()
for row in rows:
    print(json.dumps(row))
worst = max(r["ms"] for r in rows)
print(json.dumps({"full_ms": t_full, "ideal_ms": t_full / n, "worst_rank_ms": worst, "efficiency": t_full / n / worst,
                  "segment_imbalance": max(r["segments"] for r in rows) / (sum(r["segments"] for r in rows) / n)}))
