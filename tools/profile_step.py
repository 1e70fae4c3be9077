"""One C2 render (1024x1024, `spp` samples) for ncu captures; prints nothing that is a bench value."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rustlight_b200 import SceneLoaderManager, _abi  # noqa: E402
from rustlight_b200.device import Context, DeviceScene  # noqa: E402

spp = int(sys.argv[1]) if len(sys.argv) > 1 else 16
sc = SceneLoaderManager().load(os.path.join(ROOT, "data", "cbox.pbrt")).scale_image(2.0)
ctx = Context(0)
dev = DeviceScene(ctx, sc)
_, st = dev.render(_abi.path_desc(), spp, seed=0, want_image=False)
print("done", st.segments)
