"""Mixed-material Cornell box (tests/conftest.py: mixed_cbox) at 1024x1024: stage times with and without the material sort."""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import mixed_cbox  # noqa: E402
from rustlight_b200 import _abi  # noqa: E402
from rustlight_b200.device import Context, DeviceScene  # noqa: E402

spp = int(sys.argv[1]) if len(sys.argv) > 1 else 16
sc = mixed_cbox(1024, 1024)
ctx = Context(0)
dev = DeviceScene(ctx, sc)
integ = _abi.path_desc()
dev.render(integ, 2, want_image=False)
for sort in (0, 1):
    best = min(dev.render(integ, spp, want_image=False, material_sort=sort)[1].ms_total for _ in range(3))
    ctx.set_profiling(True)
    img, st = dev.render(integ, spp, material_sort=sort)
    ctx.set_profiling(False)
    print(json.dumps({"sort": sort, "ms_total": best, "trace": st.ms_trace, "shade": st.ms_shade, "shadow": st.ms_shadow, "segments": st.segments,
                      "md5": hashlib.md5(img.tobytes()).hexdigest()[:10]}))
