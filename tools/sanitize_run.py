"""Small renders that touch every kernel (path with and without material sort, direct, ao, mixed materials, textures, delta
lights, LBVH scene), for compute-sanitizer:  compute-sanitizer --tool memcheck|racecheck python tools/sanitize_run.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from conftest import load_cbox, mixed_cbox  # noqa: E402
from rustlight_b200 import SceneLoaderManager, _abi  # noqa: E402
from rustlight_b200.device import Context, DeviceScene  # noqa: E402
from rustlight_b200.host import material_diffuse  # noqa: E402
from tess_cbox import tessellated_cbox_json  # noqa: E402

ctx = Context(0)
sc = load_cbox(48, 48)
dev = DeviceScene(ctx, sc)
for integ in (_abi.path_desc(), _abi.direct_desc(2, 2), _abi.ao_desc(0.5, True)):
    dev.render(integ, 3, seed=1)
ctx.set_profiling(True)
dev.render(_abi.path_desc(), 2, seed=1)
ctx.set_profiling(False)
dev.primary_hits()
dev.close()
mx = mixed_cbox(48, 48)
mx.add_point_light((0.5, 0.5, 0.5), (0.2, 1.0, 0.3))
mx.add_directional_light((0.4, 0.4, 0.4), (0.2, -1.0, -0.1))
mx.set_environment((0.2, 0.2, 0.3))
t = mx.add_checkerboard_texture((0.8, 0.8, 0.8), (0.1, 0.1, 0.1), (0, 0), (2, 2))
mx.set_material(1, material_diffuse(kd_texture=t))
dev = DeviceScene(ctx, mx)
for sort in (0, 1):
    dev.render(_abi.path_desc(), 3, seed=2, material_sort=sort)
dev.render(_abi.direct_desc(2, 2), 2, seed=2)
dev.close()
ts = SceneLoaderManager().load_string(tessellated_cbox_json(3), "json")
ts.set_resolution(32, 32)
dev = DeviceScene(ctx, ts)
dev.render(_abi.path_desc(), 2, seed=3)
o = np.random.default_rng(0).uniform(-0.9, 0.9, (500, 3)).astype(np.float32) + np.float32([0, 1, 0])
dev.visible(o, o[::-1].copy())
dev.close()
ctx.close()
print("sanitize_run: done")
