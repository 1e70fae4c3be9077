"""Where the e2e step's extra time goes on one GPU: scene create + destroy, the frame's read-back (library stats ms_d2h and wall), and a
plain pinned D2H copy of the same size through torch for comparison (development aid)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from rustlight_b200 import SceneLoaderManager, _abi
from rustlight_b200.device import Context, DeviceScene, PinnedImage
sc = SceneLoaderManager().load(os.path.join(ROOT, "data", "cbox.pbrt")).scale_image(2.0)
ctx = Context(0)
d = DeviceScene(ctx, sc)
integ = _abi.path_desc()
pin = PinnedImage(1024, 1024)
out = pin.array
for _ in range(3):
    d.render(integ, 1, out=out)
ts, d2h = [], []
for _ in range(20):
    t0 = time.perf_counter()
    _, st = d.render(integ, 1, out=out)
    ts.append((time.perf_counter() - t0) * 1e3)
    d2h.append(st.ms_d2h)
t2 = []
for _ in range(20):
    t0 = time.perf_counter()
    d.render(integ, 1, want_image=False)
    t2.append((time.perf_counter() - t0) * 1e3)
print("render 1 spp with read-back %.3f ms (library ms_d2h %.3f), without %.3f ms; pinned out: %s" % (min(ts), min(d2h), min(t2), out is not None))
x = torch.empty(1024 * 1024 * 3, dtype=torch.float32, device="cuda")
h = torch.empty(1024 * 1024 * 3, dtype=torch.float32).pin_memory()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
best = 1e9
for _ in range(10):
    e0.record(); h.copy_(x, non_blocking=True); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print("torch pinned D2H of 12.6 MB: %.3f ms = %.1f GB/s" % (best, 12.582912 / best))
