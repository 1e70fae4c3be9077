"""One rank's eighth of config 2 (Context(0, nranks=8, rank=0), no communicator) rendered twice: for an ncu launch list of the second
frame (development aid):  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/profile_rank.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rustlight_b200 import SceneLoaderManager, _abi  # noqa: E402
from rustlight_b200.device import Context, DeviceScene  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
sc = SceneLoaderManager().load(os.path.join(ROOT, "data", "cbox.pbrt")).scale_image(2.0)
ctx = Context(0, nranks=n, rank=0)
dev = DeviceScene(ctx, sc)
integ = _abi.path_desc()
dev.render(integ, 128, want_image=False)  # (sets the queue-length prediction of the second frame)
_, st = dev.render(integ, 128, want_image=False)
print("ms_total", st.ms_total, "launches", st.kernel_launches)
