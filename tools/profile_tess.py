"""One render of the tessellated Cornell box (tools/tess_cbox.py N, 1024x1024, `spp` samples) for ncu captures of the tree kernels."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from rustlight_b200 import SceneLoaderManager, _abi  # noqa: E402
from rustlight_b200.device import Context, DeviceScene  # noqa: E402
from tess_cbox import tessellated_cbox_json  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 4
sc = SceneLoaderManager().load_string(tessellated_cbox_json(n), "json").scale_image(2.0)
ctx = Context(0)
dev = DeviceScene(ctx, sc)
_, st = dev.render(_abi.path_desc(), spp, seed=0, want_image=False)
print("done", st.segments)
