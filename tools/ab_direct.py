"""`direct` (config C4 shape: 2048x2048, -b 1 -l 1) and `ao` stage times on one GPU (development aid)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rustlight_b200 import SceneLoaderManager, _abi  # noqa: E402
from rustlight_b200.device import Context, DeviceScene  # noqa: E402

spp = int(sys.argv[1]) if len(sys.argv) > 1 else 16
sc = SceneLoaderManager().load(os.path.join(ROOT, "data", "cbox.pbrt")).scale_image(4.0)
ctx = Context(0)
dev = DeviceScene(ctx, sc)
for name, integ in (("direct b1 l1", _abi.direct_desc(1, 1)), ("direct b2 l2", _abi.direct_desc(2, 2)), ("ao", _abi.ao_desc(1.0, False))):
    dev.render(integ, 2, want_image=False)
    best = min(dev.render(integ, spp, want_image=False)[1].ms_total for _ in range(3))
    ctx.set_profiling(True)
    _, st = dev.render(integ, spp, want_image=False)
    ctx.set_profiling(False)
    print(json.dumps({"integrator": name, "spp": spp, "ms_total": best, "Msamples/s": st.samples / best / 1e3, "Mrays/s": (st.segments + st.shadow_rays) / best / 1e3,
                      "trace": st.ms_trace, "shade": st.ms_shade, "shadow": st.ms_shadow, "raygen": st.ms_raygen, "accum": st.ms_accum}))
