"""A/B harness for kernel variants built into build/variants/librl_<name>.so (development aid).
Each variant runs in its own process (RL_B200_LIB), renders cbox 1024x1024 x SPP with per-stage
CUDA-event timing, and prints stage times + an image checksum (all variants must agree)."""
import glob
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SPP = int(os.environ.get("AB_SPP", "32"))

CHILD = r'''
import os, sys, json, hashlib
sys.path.insert(0, %r)
from rustlight_b200 import SceneLoaderManager, _abi
from rustlight_b200.device import Context, DeviceScene
sc = SceneLoaderManager().load(os.path.join(%r, "data", "cbox.pbrt")).scale_image(2.0)
ctx = Context(0); dev = DeviceScene(ctx, sc); integ = _abi.path_desc()
dev.render(integ, 4, want_image=False)
best = None
for _ in range(3):
    _, st = dev.render(integ, %d, want_image=False)
    best = st.ms_total if best is None else min(best, st.ms_total)
ctx.set_profiling(True)
img, st = dev.render(integ, %d)
print(json.dumps({"ms_total": best, "trace": st.ms_trace, "shade": st.ms_shade, "shadow": st.ms_shadow, "raygen": st.ms_raygen,
                  "segments": st.segments, "md5": hashlib.md5(img.tobytes()).hexdigest()[:10]}))
''' % (ROOT, ROOT, SPP, SPP)

libs = sys.argv[1:] or sorted(glob.glob(os.path.join(ROOT, "build", "variants", "librl_*.so")))
for lib in libs:
    env = dict(os.environ, RL_B200_LIB=lib)
    out = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    line = out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-400:]
    print(f"{os.path.basename(lib):28s} {line}", flush=True)
