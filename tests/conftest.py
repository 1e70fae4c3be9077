import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

DATA = os.path.join(ROOT, "data")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _ensure_built():
    """(Re)build every native piece whose sources changed: CUDA library, host layer, oracle, emulator."""
    import __graft_entry__ as g
    g.build()


_ensure_built()


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container (GPU tests run under gpurun)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def loader():
    from rustlight_b200 import SceneLoaderManager
    return SceneLoaderManager()


def load_cbox(w=None, h=None, name="cbox.pbrt"):
    from rustlight_b200 import SceneLoaderManager
    sc = SceneLoaderManager().load(os.path.join(DATA, name))
    if w is not None:
        sc.set_resolution(w, h or w)
    return sc


@pytest.fixture(scope="session")
def cbox():
    return load_cbox()


@pytest.fixture(scope="session")
def cbox_oracle(cbox):
    from oracle import binding as ob
    return ob.OracleScene(cbox)


@pytest.fixture(scope="session")
def cbox64():
    return load_cbox(64, 64)


@pytest.fixture(scope="session")
def gpu_ctx():
    from rustlight_b200.device import Context
    ctx = Context(0)
    yield ctx
    ctx.close()


def soup_scene(ntris, seed=0, w=64, h=64, emissive_first=True):
    """Random triangle soup inside the unit cube + one emissive quad, as scene JSON text."""
    import json
    rng = np.random.default_rng(seed)
    meshes = []
    per = 50
    nm = max(1, ntris // per)
    for m in range(nm):
        c = rng.uniform(-0.8, 0.8, (per, 1, 3))
        tri = c + rng.normal(scale=0.15, size=(per, 3, 3))
        P = tri.reshape(-1, 3).astype(np.float32)
        idx = np.arange(per * 3)
        kd = rng.uniform(0.2, 0.8, 3)
        meshes.append({"name": f"soup{m}", "material": {"type": "diffuse", "kd": [float(x) for x in kd]},
                       "indices": [int(i) for i in idx], "P": [float(x) for x in P.ravel()]})
    if emissive_first:
        meshes.append({"name": "light", "material": {"type": "diffuse", "kd": [0, 0, 0]}, "emission": [10, 10, 10],
                       "indices": [0, 1, 2, 0, 2, 3],
                       "P": [-0.5, 0.95, -0.5, 0.5, 0.95, -0.5, 0.5, 0.95, 0.5, -0.5, 0.95, 0.5],
                       "N": [0, -1, 0] * 4})
    cam = {"width": w, "height": h, "fov": 40.0, "fov_axis": "y", "flip": False,
           "to_world": [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, -1, 0, 0, 0, 3.5, 1]}
    return json.dumps({"camera": cam, "meshes": meshes})


def adversarial_case(seed):
    """Scene + rays that stress culling and the leaf prefilter: tiny / huge / sliver / far-away triangles,
    rays aimed at vertices and edges, grazing directions, origins from 1e-3 to 1e2 scene units away."""
    import json
    from rustlight_b200 import SceneLoaderManager
    rng = np.random.default_rng(100 + seed)
    scale = [1e-3, 1.0, 50.0, 1.0, 1.0, 1e3][seed]
    shift = np.array([[0, 0, 0], [0, 0, 0], [0, 0, 0], [300, -200, 100], [0, 0, 0], [0, 0, 0]][seed], np.float64)
    meshes = []
    for m in range(6):
        n = 8
        c = rng.uniform(-1, 1, (n, 1, 3)) * scale + shift
        ext = scale * 10.0 ** rng.uniform(-3, 0, (n, 1, 1))
        tri = c + rng.normal(size=(n, 3, 3)) * ext
        if m % 2 == 0:  # slivers: third vertex almost on the first edge
            tri[:, 2] = tri[:, 0] + (tri[:, 1] - tri[:, 0]) * rng.uniform(0.2, 0.8, (n, 1)) + rng.normal(size=(n, 3)) * ext[:, 0] * 1e-3
        meshes.append({"material": {"type": "diffuse", "kd": [0.5] * 3}, "indices": list(range(3 * n)),
                       "P": [float(x) for x in tri.astype(np.float32).ravel()]})
    meshes[0]["emission"] = [1, 1, 1]
    txt = json.dumps({"camera": {"width": 8, "height": 8, "fov": 40, "to_world": [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, -1, 0, 0, 0, 3, 1]}, "meshes": meshes})
    sc = SceneLoaderManager().load_string(txt, "json")
    d = sc.desc.contents
    verts = np.concatenate([np.ctypeslib.as_array(d.meshes[i].P, (3 * d.meshes[i].nverts,)).reshape(-1, 3) for i in range(d.nmeshes)])
    tris = verts.reshape(-1, 3, 3).astype(np.float64)
    nr = 4000
    pick = rng.integers(0, len(tris), nr)
    bary = rng.dirichlet([0.3, 0.3, 0.3], nr)                      # clustered toward edges and vertices
    bary[::5] = np.eye(3)[rng.integers(0, 3, len(bary[::5]))]      # exactly at a vertex
    bary[1::5, 2] = 0.0
    bary[1::5, :2] = rng.dirichlet([1, 1], len(bary[1::5]))        # exactly on an edge
    tgt = (tris[pick] * bary[:, :, None]).sum(axis=1)
    nrm = np.cross(tris[pick, 1] - tris[pick, 0], tris[pick, 2] - tris[pick, 0])
    nrm /= np.maximum(np.linalg.norm(nrm, axis=1, keepdims=True), 1e-30)
    dirs = rng.normal(size=(nr, 3))
    g = dirs[::2]
    dirs[::2] = g - (g * nrm[::2]).sum(axis=1, keepdims=True) * nrm[::2] * (1 - 10.0 ** rng.uniform(-6, -1, (len(g), 1)))  # grazing
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    dist = scale * 10.0 ** rng.uniform(-3, 2, (nr, 1))
    o = (tgt - dirs * dist).astype(np.float32)
    dd = dirs.astype(np.float32)
    dd /= np.linalg.norm(dd.astype(np.float64), axis=1, keepdims=True).astype(np.float32)
    p1 = (tgt + rng.normal(size=(nr, 3)) * scale * 1e-2).astype(np.float32)
    return sc, o, dd, p1


def adversarial_pairs_case(seed):
    """Scene of quads (two triangles each) that stresses the PAIR records of the flat group table: the fourth
    vertex sits 1e-9..3e-6 quad sizes off the plane of the first triangle (straddling the pairing thresholds:
    normals equal to 4e-7, mismatch below 1e-6 * abs_max), quads are parallelograms, kites or slivers, the second triangle is wound either way;
    rays aim at the shared diagonal, at edges and vertices, half of them grazing."""
    import json
    from rustlight_b200 import SceneLoaderManager
    rng = np.random.default_rng(500 + seed)
    scale = [1.0, 1e-3, 200.0, 1.0][seed]
    shift = np.array([[0, 0, 0], [0, 0, 0], [0, 0, 0], [40, -25, 10]][seed], np.float64)
    nq = [14, 12, 10, 14][seed]
    meshes, quads = [], []
    for q in range(nq):
        c = rng.uniform(-1, 1, 3) * scale + shift
        e1 = rng.normal(size=3)
        e1 /= np.linalg.norm(e1)
        e2 = np.cross(e1, rng.normal(size=3))
        e2 /= np.linalg.norm(e2)
        nrm = np.cross(e1, e2)
        if q % 4 == 0:  # axis-aligned face
            e1, e2, nrm = np.eye(3)[[q % 3, (q + 1) % 3, (q + 2) % 3]]
        s1, s2 = scale * 10.0 ** rng.uniform(-2, 0, 2)
        if q % 5 == 1:
            s2 = s1 * 1e-3  # sliver quad
        a, b, d = c, c + e1 * s1, c + e2 * s2
        cc = c + e1 * s1 * rng.uniform(0.6, 1.4) + e2 * s2 * rng.uniform(0.6, 1.4) if q % 2 else c + e1 * s1 + e2 * s2
        cc = cc + nrm * min(s1, s2) * 10.0 ** rng.uniform(-9, -5.5) * rng.choice([-1, 1])
        P = np.array([a, b, cc, d], np.float32)
        idx = [0, 1, 2, 0, 2, 3] if q % 3 else [0, 1, 2, 0, 3, 2]
        quads.append(P.astype(np.float64))
        meshes.append({"material": {"type": "diffuse", "kd": [0.5] * 3}, "indices": idx, "P": [float(x) for x in P.ravel()]})
    meshes[0]["emission"] = [1, 1, 1]
    txt = json.dumps({"camera": {"width": 8, "height": 8, "fov": 40, "to_world": [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, -1, 0, 0, 0, 3, 1]}, "meshes": meshes})
    sc = SceneLoaderManager().load_string(txt, "json")
    nr = 6000
    pick = rng.integers(0, nq, nr)
    Q = np.array(quads)[pick]                                     # nr x 4 x 3
    w = rng.dirichlet([0.4, 0.4, 0.4, 0.4], nr)
    w[::6] = 0.0
    w[::6, 0] = rng.uniform(0, 1, len(w[::6]))                    # on the shared diagonal (vertices 0 and 2)
    w[::6, 2] = 1.0 - w[::6, 0]
    w[1::6] = np.eye(4)[rng.integers(0, 4, len(w[1::6]))]         # exactly at a vertex
    k = rng.integers(0, 4, len(w[2::6]))                          # on an outer edge
    t = rng.uniform(0, 1, len(k))
    w[2::6] = np.eye(4)[k] * t[:, None] + np.eye(4)[(k + 1) % 4] * (1 - t[:, None])
    tgt = (Q * w[:, :, None]).sum(axis=1)
    nrm = np.cross(Q[:, 1] - Q[:, 0], Q[:, 3] - Q[:, 0])
    nrm /= np.maximum(np.linalg.norm(nrm, axis=1, keepdims=True), 1e-300)
    dirs = rng.normal(size=(nr, 3))
    g = dirs[::2]
    dirs[::2] = g - (g * nrm[::2]).sum(axis=1, keepdims=True) * nrm[::2] * (1 - 10.0 ** rng.uniform(-7, -1, (len(g), 1)))  # grazing
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    dist = scale * 10.0 ** rng.uniform(-3, 2, (nr, 1))
    o = (tgt - dirs * dist).astype(np.float32)
    dd = dirs.astype(np.float32)
    dd /= np.linalg.norm(dd.astype(np.float64), axis=1, keepdims=True).astype(np.float32)
    p1 = (tgt + rng.normal(size=(nr, 3)) * scale * 1e-2).astype(np.float32)
    return sc, o, dd, p1


def mixed_cbox(w=48, h=48):
    """Cornell box with every material kind: floor substrate (GGX), ceiling phong, back wall rough gold (GGX), right wall
    mirror, left wall substrate without a distribution (delta + diffuse), short box glass (closed mesh), tall box
    Beckmann aluminium."""
    from rustlight_b200.host import material_glass, material_metal, material_mirror, material_phong, material_substrate
    sc = load_cbox(w, h)
    sc.set_material(0, material_substrate((0.4, 0.25, 0.1), (0.05, 0.05, 0.05), "ggx", 0.1))
    sc.set_material(1, material_phong((0.4, 0.4, 0.4), (0.3, 0.3, 0.3), 40.0))
    sc.set_material(2, material_metal((1, 1, 1), (0.143, 0.375, 1.442), (3.983, 2.386, 1.603), "ggx", 0.12))
    sc.set_material(3, material_mirror((0.9, 0.8, 0.7)))
    sc.set_material(4, material_substrate((0.5, 0.5, 0.5), (0.04, 0.04, 0.04), None, 0.0))
    sc.set_material(5, material_glass((1, 1, 1), (0.95, 0.97, 1.0), 1.5046, 1.000277))
    sc.set_material(6, material_metal((0.9, 0.9, 0.9), (1.657, 0.880, 0.521), (9.224, 6.270, 4.837), "beckmann", 0.25))
    return sc


def rel_l2(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b.astype(np.float64)) / max(np.linalg.norm(b.astype(np.float64)), 1e-30))


def author_metrics(ref, test, eps=1e-2):
    """The image metrics the reference's author uses (tests/interactive-viewer/tools/metric.py:18-45, compute_metric):
    means over pixels and channels of l1, l2, mrse, mape, smape with eps = 1e-2."""
    ref, test = ref.astype(np.float64), test.astype(np.float64)
    diff = ref - test
    return {"l1": float(np.abs(diff).mean()), "l2": float((diff * diff).mean()), "mrse": float((diff * diff / (ref * ref + eps)).mean()),
            "mape": float((np.abs(diff) / (ref + eps)).mean()), "smape": float((2 * np.abs(diff) / (ref + test + eps)).mean())}


def blend_triple(a, b, weight):
    """One BSDFBlend as the per-material test entry points of the oracle and the emulator take it: the first element of an array
    {blend, bsdf1, bsdf2} with blend_a = 1, blend_b = 2 (the returned struct is a view into that array and keeps it alive)."""
    import ctypes as C
    from rustlight_b200 import _abi
    arr = (_abi.rl_material * 3)()
    C.memmove(C.byref(arr, C.sizeof(_abi.rl_material)), C.byref(a), C.sizeof(_abi.rl_material))
    C.memmove(C.byref(arr, 2 * C.sizeof(_abi.rl_material)), C.byref(b), C.sizeof(_abi.rl_material))
    arr[0].kind = _abi.RL_BSDF_BLEND
    arr[0].blend_a, arr[0].blend_b, arr[0].blend_weight = 1, 2, weight
    return arr[0]


def blended_cbox(w=48, h=48):
    """Cornell box with BSDFBlend on four meshes (every rough kind appears as a part), the other meshes as in the file."""
    from rustlight_b200.host import material_metal, material_phong, material_substrate
    sc = load_cbox(w, h)
    diffuse = _diffuse((0.6, 0.5, 0.2))
    sc.set_material_blend(0, diffuse, material_phong((0.2, 0.2, 0.2), (0.5, 0.5, 0.5), 30.0), 0.35)
    sc.set_material_blend(2, material_metal((1, 1, 1), (0.143, 0.375, 1.442), (3.983, 2.386, 1.603), "ggx", 0.2), _diffuse((0.1, 0.3, 0.6)), 0.5)
    sc.set_material_blend(5, material_substrate((0.4, 0.25, 0.1), (0.05, 0.05, 0.05), "beckmann", 0.3), material_phong((0.1, 0.1, 0.1), (0.6, 0.6, 0.6), 80.0), 0.75)
    sc.set_material_blend(6, _diffuse((0.7, 0.7, 0.7)), _diffuse((0.1, 0.6, 0.1)), 0.25)
    return sc


def _diffuse(kd):
    from rustlight_b200 import _abi
    m = _abi.rl_material()
    m.kind = _abi.RL_BSDF_DIFFUSE
    m.kd[0], m.kd[1], m.kd[2] = kd
    m.ior = 1.0
    return m
