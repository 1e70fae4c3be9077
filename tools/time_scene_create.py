"""How long do rl_scene_create + rl_scene_destroy take (the per-step scene rebuild inside bench.py's e2e region)?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rustlight_b200 import SceneLoaderManager, _abi
from rustlight_b200.device import Context, DeviceScene
sc = SceneLoaderManager().load(os.path.join(ROOT, "data", "cbox.pbrt")).scale_image(2.0)
ctx = Context(0)
for _ in range(3):
    DeviceScene(ctx, sc).close()
t0 = time.perf_counter()
n = 50
for _ in range(n):
    d = DeviceScene(ctx, sc)
    d.close()
print("scene create + destroy: %.3f ms" % ((time.perf_counter() - t0) / n * 1e3))
d = DeviceScene(ctx, sc)
integ = _abi.path_desc()
d.render(integ, 1)
t0 = time.perf_counter()
for _ in range(20):
    d.render(integ, 1, want_image=True)
print("render 1 spp incl. read-back: %.3f ms" % ((time.perf_counter() - t0) / 20 * 1e3))
t0 = time.perf_counter()
for _ in range(20):
    d.render(integ, 1, want_image=False)
print("render 1 spp without read-back: %.3f ms" % ((time.perf_counter() - t0) / 20 * 1e3))
