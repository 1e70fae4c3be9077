"""Timings of the other BASELINE.json configs on one GPU (parity-test cases, not bench lines): C1 path 512^2 x 16 spp,
C3 Phong walls path 512^2 x 512 spp, C4 direct 2048^2 x 64 spp, C5 path 1920x1080 (spp argument, default 64 of 4096)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_cbox  # noqa: E402
from rustlight_b200 import _abi  # noqa: E402
from rustlight_b200.device import Context, DeviceScene  # noqa: E402
from rustlight_b200.host import material_phong  # noqa: E402

c5_spp = int(sys.argv[1]) if len(sys.argv) > 1 else 64
ctx = Context(0)


def run(name, sc, integ, spp, **kw):
    dev = DeviceScene(ctx, sc)
    dev.render(integ, min(spp, 4), want_image=False, **kw)
    best = min((dev.render(integ, spp, want_image=False, **kw)[1] for _ in range(3)), key=lambda s: s.ms_total)
    print(json.dumps({"config": name, "spp": spp, "ms": best.ms_total, "Msamples/s": best.samples / best.ms_total / 1e3,
                      "Msegments/s": best.segments / best.ms_total / 1e3, "Mshadow_rays/s": best.shadow_rays / best.ms_total / 1e3,
                      "launches": best.kernel_launches}))
    dev.close()


run("C1 cbox path 512x512 x16", load_cbox(512, 512), _abi.path_desc(), 16)
sc = load_cbox(512, 512)
kds = [(0.63, 0.065, 0.05), (0.14, 0.45, 0.091), (0.725, 0.71, 0.68)]
for mesh, kd in [(0, kds[2]), (1, kds[2]), (2, kds[2]), (3, kds[1]), (4, kds[0])]:
    sc.set_material(mesh, material_phong([0.5 * c for c in kd], (0.3, 0.3, 0.3), 50.0))
run("C3 Phong walls path 512x512 x512", sc, _abi.path_desc(), 512)
run("C4 cbox direct -b 1 -l 1 2048x2048 x64", load_cbox(2048, 2048), _abi.direct_desc(1, 1), 64)
run(f"C5 cbox path 1920x1080 x{c5_spp} (of 4096), material sort on", load_cbox(1920, 1080), _abi.path_desc(), c5_spp, material_sort=1)
