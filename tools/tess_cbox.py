"""Cornell box with every quad face tessellated into n x n cells (2 n^2 triangles per face): the same picture through the
LBVH traversal kernels instead of the group table.  Used by tests (image ~ flat Cornell box) and for LBVH timings.

usage as a script: python tools/tess_cbox.py N [spp]  -> stage times of the tessellated and the plain scene on one GPU"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def tessellated_cbox_json(n):
    from rustlight_b200 import SceneLoaderManager
    base = json.loads(SceneLoaderManager().load(os.path.join(ROOT, "data", "cbox.pbrt")).to_json())
    meshes = []
    for m in base["meshes"]:
        P = np.array(m["P"], np.float64).reshape(-1, 3)
        N = np.array(m["N"], np.float64).reshape(-1, 3) if "N" in m else None
        idx = np.array(m["indices"]).reshape(-1, 3)
        newP, newN, newI = [], [], []
        for t0 in range(0, len(idx), 2):  # two triangles = one quad (0,1,2),(0,2,3) or the box winding (0,2,1),(0,3,2)
            a, b, c = idx[t0]
            d = [v for v in idx[t0 + 1] if v not in (a, b, c)][0]
            quad = [a, b, c, d] if idx[t0 + 1][1] == c else [a, c, b, d]
            # corners in order around the quad: q0 q1 q2 q3 with triangles (q0,q1,q2),(q0,q2,q3) up to winding
            q = [P[i] for i in (quad[0], quad[1], quad[2], quad[3])]
            nq = [N[i] for i in quad] if N is not None else None
            flip = idx[t0 + 1][1] != c
            base_i = len(newP)
            for j in range(n + 1):
                for i in range(n + 1):
                    u, v = i / n, j / n
                    w = [(1 - u) * (1 - v), u * (1 - v), u * v, (1 - u) * v]
                    newP.append(sum(wk * qk for wk, qk in zip(w, q)))
                    if nq is not None:
                        newN.append(sum(wk * nk for wk, nk in zip(w, nq)))
            for j in range(n):
                for i in range(n):
                    p00 = base_i + j * (n + 1) + i
                    p10, p01, p11 = p00 + 1, p00 + n + 1, p00 + n + 2
                    tris = [(p00, p10, p11), (p00, p11, p01)]
                    if flip:
                        tris = [(x, z, y) for x, y, z in tris]
                    newI += tris
        mm = dict(m)
        mm["P"] = [float(x) for x in np.array(newP, np.float32).ravel()]
        mm["indices"] = [int(x) for x in np.array(newI).ravel()]
        if N is not None:
            mm["N"] = [float(x) for x in np.array(newN, np.float32).ravel()]
        mm.pop("uv", None)
        meshes.append(mm)
    base["meshes"] = meshes
    return json.dumps(base)


def main():
    from rustlight_b200 import SceneLoaderManager, _abi
    from rustlight_b200.device import Context, DeviceScene
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    spp = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    ctx = Context(0)
    integ = _abi.path_desc()
    for name, sc in (("plain", SceneLoaderManager().load(os.path.join(ROOT, "data", "cbox.pbrt"))),
                     (f"tess{n}", SceneLoaderManager().load_string(tessellated_cbox_json(n), "json"))):
        sc.scale_image(2.0)
        dev = DeviceScene(ctx, sc)
        bi = dev.bvh_info()
        dev.render(integ, 2, want_image=False)
        best = min(dev.render(integ, spp, want_image=False)[1].ms_total for _ in range(3))
        ctx.set_profiling(True)
        img, st = dev.render(integ, spp)
        ctx.set_profiling(False)
        print(json.dumps({"scene": name, "ntris": bi.ntris, "bvh_depth": bi.max_depth, "smem_resident": bi.smem_resident, "flat_groups": bi.flat_groups,
                          "ms_total": best, "Mseg/s": st.segments / best / 1e3, "trace": st.ms_trace, "shade": st.ms_shade, "shadow": st.ms_shadow,
                          "mean": float(img.mean()), "md5": __import__("hashlib").md5(img.tobytes()).hexdigest()[:10], "launches": st.kernel_launches}))
        dev.close()


if __name__ == "__main__":
    main()
