"""Material sort on/off timing on the diffuse box (config 5 shape) and the Phong-walls box (config 3)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rustlight_b200 import SceneLoaderManager, _abi
from rustlight_b200.device import Context, DeviceScene
from rustlight_b200.host import material_phong
ctx = Context(0)
for name in ("diffuse", "phong"):
    sc = SceneLoaderManager().load(os.path.join(ROOT, "data", "cbox.pbrt")).scale_image(2.0)
    if name == "phong":
        kds = [(0.63, 0.065, 0.05), (0.14, 0.45, 0.091), (0.725, 0.71, 0.68)]
        for mesh, kd in [(0, kds[2]), (1, kds[2]), (2, kds[2]), (3, kds[1]), (4, kds[0])]:
            sc.set_material(mesh, material_phong([0.5 * c for c in kd], (0.3, 0.3, 0.3), 50.0))
    dev = DeviceScene(ctx, sc)
    integ = _abi.path_desc()
    dev.render(integ, 4, want_image=False)
    for ms in (0, 1):
        best = min(dev.render(integ, 32, want_image=False, material_sort=ms)[1].ms_total for _ in range(3))
        ctx.set_profiling(True); _, st = dev.render(integ, 32, want_image=False, material_sort=ms); ctx.set_profiling(False)
        print(json.dumps({"scene": name, "material_sort": ms, "ms_total": best, "shade": st.ms_shade, "trace": st.ms_trace, "shadow": st.ms_shadow,
                          "Msamples/s": st.samples / best / 1e3, "segments": st.segments}))
    dev.close()
