"""Host layer (C++ loaders, camera, PFM) and the C-ABI surface; no GPU needed."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

from conftest import DATA, ROOT, load_cbox
from rustlight_b200 import SceneError, SceneLoaderManager, _abi
from rustlight_b200.host import read_pfm, save_pfm


def test_cbox_pbrt_contents(cbox):
    d = cbox.desc.contents
    assert d.nmeshes == 8 and cbox.nb_triangles == 36 and cbox.size == (512, 512)
    assert [d.meshes[i].ntris for i in range(8)] == [2, 2, 2, 2, 2, 12, 12, 2]
    assert [d.meshes[i].emission_kind for i in range(8)] == [0] * 7 + [1]
    assert list(d.meshes[7].emission) == [17.0, 12.0, 4.0]
    assert np.allclose(list(d.meshes[4].mat.kd), [0.63, 0.065, 0.05])   # LeftWall: red
    assert np.allclose(list(d.meshes[3].mat.kd), [0.14, 0.45, 0.091])   # RightWall: green
    assert all(d.meshes[i].mat.kind == _abi.RL_BSDF_DIFFUSE for i in range(8))
    assert d.meshes[0].N and d.meshes[0].UV
    # light area 0.47 x 0.38 = 0.1786
    P = np.ctypeslib.as_array(d.meshes[7].P, (12,)).reshape(4, 3)
    assert np.ptp(P[:, 0]) * np.ptp(P[:, 2]) == pytest.approx(0.1786, rel=1e-6)


def test_json_is_a_faithful_transcription(cbox):
    js = SceneLoaderManager().load(os.path.join(DATA, "cbox.json"))
    a, b = cbox.desc.contents, js.desc.contents
    assert a.nmeshes == b.nmeshes
    assert bytes(a.camera) == bytes(b.camera)
    for i in range(a.nmeshes):
        ma, mb = a.meshes[i], b.meshes[i]
        assert (ma.nverts, ma.ntris, ma.emission_kind) == (mb.nverts, mb.ntris, mb.emission_kind)
        for f, n in (("P", 3 * ma.nverts), ("N", 3 * ma.nverts), ("UV", 2 * ma.nverts)):
            assert np.array_equal(np.ctypeslib.as_array(getattr(ma, f), (n,)), np.ctypeslib.as_array(getattr(mb, f), (n,)))
        assert np.array_equal(np.ctypeslib.as_array(ma.idx, (3 * ma.ntris,)), np.ctypeslib.as_array(mb.idx, (3 * ma.ntris,)))
        assert bytes(ma.mat) == bytes(mb.mat) and list(ma.emission) == list(mb.emission)
    # and the writer round-trips
    again = SceneLoaderManager().load_string(js.to_json(), "json")
    assert bytes(again.desc.contents.camera) == bytes(a.camera)


def test_use_shading_normal_flag():
    sc = SceneLoaderManager().load(os.path.join(DATA, "cbox.pbrt"), use_shading_normal=False)
    assert not sc.desc.contents.meshes[0].N  # `-x no-shading`: normals dropped (scene_loader.rs:101-118)


def test_loader_errors():
    m = SceneLoaderManager()
    with pytest.raises(SceneError, match="extension"):
        m.load("scene.obj")            # no such loader registered (scene_loader.rs:40-43)
    with pytest.raises(SceneError, match="No file extension"):
        m.load("scene")
    with pytest.raises(SceneError, match="camera"):
        m.load_string('WorldBegin Shape "trianglemesh" "integer indices" [0 1 2] "point P" [0 0 0 1 0 0 0 1 0] WorldEnd', "pbrt")
    with pytest.raises(SceneError, match="not supported"):
        m.load_string('Camera "perspective" WorldBegin MakeNamedMaterial "g" "string type" ["uber"] WorldEnd', "pbrt")
    with pytest.raises(SceneError, match="anisotropic"):  # distribution.rs:62 asserts alpha_u == alpha_v
        m.load_string('Camera "perspective" WorldBegin Material "metal" "float uroughness" [0.1] "float vroughness" [0.2] WorldEnd', "pbrt")
    with pytest.raises(SceneError, match="out of range"):
        m.load_string('Camera "perspective" WorldBegin Shape "trianglemesh" "integer indices" [0 1 5] "point P" [0 0 0 1 0 0 0 1 0] WorldEnd', "pbrt")


def test_pbrt_material_mapping():
    """bsdf_pbrt (bsdfs/mod.rs:294-386): mirror -> BSDFMetal without a distribution, metal / substrate -> GGX with the
    remapped roughness (distribution_pbrt, :260-291), glass -> BSDFGlass.eta(eta, 1.0), distribution ignored."""
    import math
    from rustlight_b200.host import remap_roughness
    tri = 'Shape "trianglemesh" "integer indices" [0 1 2] "point P" [0 0 0 1 0 0 0 1 0]'
    txt = f'''Camera "perspective" WorldBegin
      Material "mirror" "rgb Kr" [0.8 0.7 0.6] {tri}
      Material "metal" "rgb eta" [0.2 0.9 1.1] "rgb k" [3.9 2.4 2.1] "float roughness" [0.05] {tri}
      Material "metal" "float roughness" [0.3] "bool remaproughness" ["false"] {tri}
      Material "glass" "float eta" [1.33] "rgb Kt" [0.9 0.9 1.0] "float uroughness" [0.2] "float vroughness" [0.2] {tri}
      Material "substrate" "rgb Kd" [0.3 0.2 0.1] "rgb Ks" [0.04 0.04 0.04] "float uroughness" [0.02] "float vroughness" [0.02] {tri}
      Material "substrate" {tri}
    WorldEnd'''
    sc = SceneLoaderManager().load_string(txt, "pbrt")
    d = sc.desc.contents
    mats = [d.meshes[i].mat for i in range(6)]
    assert [m.kind for m in mats] == [_abi.RL_BSDF_METAL, _abi.RL_BSDF_METAL, _abi.RL_BSDF_METAL, _abi.RL_BSDF_GLASS, _abi.RL_BSDF_SUBSTRATE, _abi.RL_BSDF_SUBSTRATE]
    assert mats[0].microfacet == _abi.RL_MICROFACET_NONE and list(mats[0].ks) == pytest.approx([0.8, 0.7, 0.6]) and list(mats[0].eta) == [1, 1, 1] and list(mats[0].k) == [0, 0, 0]
    x = math.log(0.05)
    want = 1.62142 + 0.819955 * x + 0.1734 * x * x + 0.0171201 * x ** 3 + 0.000640711 * x ** 4
    assert mats[1].microfacet == _abi.RL_MICROFACET_GGX and mats[1].alpha == pytest.approx(want, rel=1e-5) == pytest.approx(remap_roughness(0.05), rel=1e-6)
    assert list(mats[1].ks) == [1, 1, 1] and list(mats[1].eta) == pytest.approx([0.2, 0.9, 1.1])
    assert mats[2].alpha == pytest.approx(0.3) and remap_roughness(1e-9) == pytest.approx(remap_roughness(1e-3))  # v.max(1e-3)
    assert mats[3].ior == pytest.approx(1.33) and list(mats[3].kt) == pytest.approx([0.9, 0.9, 1.0]) and list(mats[3].ks) == [1, 1, 1]
    assert mats[4].alpha == pytest.approx(remap_roughness(0.02)) and list(mats[4].kd) == pytest.approx([0.3, 0.2, 0.1])
    assert list(mats[5].kd) == [0.5] * 3 and list(mats[5].ks) == [0.5] * 3 and mats[5].alpha == pytest.approx(remap_roughness(0.1))


def test_pbrt_objects_instances_and_plymesh(tmp_path):
    """ObjectBegin / ObjectInstance (scene_loader.rs:183-203: instance.matrix * shape.matrix, instances after the top-level
    shapes) and Shape "plymesh" (ascii and binary_little_endian, quads fanned) relative to the scene file."""
    import struct
    (tmp_path / "q.ply").write_text("ply\nformat ascii 1.0\nelement vertex 4\nproperty float x\nproperty float y\nproperty float z\n"
                                    "property float nx\nproperty float ny\nproperty float nz\nelement face 1\n"
                                    "property list uchar int vertex_indices\nend_header\n"
                                    "0 0 0 0 0 1\n1 0 0 0 0 1\n1 1 0 0 0 1\n0 1 0 0 0 1\n4 0 1 2 3\n")
    hdr = ("ply\nformat binary_little_endian 1.0\nelement vertex 3\nproperty float x\nproperty float y\nproperty float z\n"
           "property float u\nproperty float v\nelement face 1\nproperty list uchar uint vertex_indices\nend_header\n").encode()
    body = b"".join(struct.pack("<5f", *v) for v in [(0, 0, 0, 0, 0), (2, 0, 0, 1, 0), (0, 2, 0, 0, 1)]) + struct.pack("<B3I", 3, 0, 1, 2)
    (tmp_path / "t.ply").write_bytes(hdr + body)
    (tmp_path / "s.pbrt").write_text('''Camera "perspective" WorldBegin
      ObjectBegin "quad"
        Translate 0 0 1
        Shape "plymesh" "string filename" "q.ply"
      ObjectEnd
      Shape "plymesh" "string filename" "t.ply"
      AttributeBegin Translate 10 0 0 ObjectInstance "quad" AttributeEnd
      AttributeBegin Scale 2 2 2 Material "mirror" ObjectInstance "quad" AttributeEnd
    WorldEnd''')
    sc = SceneLoaderManager().load(str(tmp_path / "s.pbrt"))
    d = sc.desc.contents
    assert d.nmeshes == 3
    P = [np.ctypeslib.as_array(d.meshes[i].P, (3 * d.meshes[i].nverts,)).reshape(-1, 3) for i in range(3)]
    idx = [list(np.ctypeslib.as_array(d.meshes[i].idx, (3 * d.meshes[i].ntris,))) for i in range(3)]
    assert idx[0] == [0, 1, 2] and np.array_equal(P[0], [[0, 0, 0], [2, 0, 0], [0, 2, 0]]) and bool(d.meshes[0].UV) and not d.meshes[0].N
    assert idx[1] == [0, 1, 2, 0, 2, 3] and np.array_equal(P[1], [[10, 0, 1], [11, 0, 1], [11, 1, 1], [10, 1, 1]])
    assert np.array_equal(P[2], [[0, 0, 2], [2, 0, 2], [2, 2, 2], [0, 2, 2]])
    assert bool(d.meshes[1].N) and list(np.ctypeslib.as_array(d.meshes[2].N, (12,))[:3]) == [0, 0, 1]  # renormalised after the scale
    assert d.meshes[1].mat.kind == d.meshes[2].mat.kind == _abi.RL_BSDF_DIFFUSE  # material bound at definition, not at instantiation
    (tmp_path / "geo.pbrt").write_text('Shape "plymesh" "string filename" "t.ply"  # included geometry\n')
    (tmp_path / "top.pbrt").write_text('Camera "perspective"\nWorldBegin\n  Include "geo.pbrt"\n  Translate 0 0 3\n  Include "geo.pbrt"\nWorldEnd\n')
    inc = SceneLoaderManager().load(str(tmp_path / "top.pbrt"))
    di = inc.desc.contents
    assert di.nmeshes == 2 and np.ctypeslib.as_array(di.meshes[1].P, (9,))[2] == 3.0
    # binary_big_endian: the same triangle
    hdr_be = hdr.replace(b"binary_little_endian", b"binary_big_endian")
    body_be = b"".join(struct.pack(">5f", *v) for v in [(0, 0, 0, 0, 0), (2, 0, 0, 1, 0), (0, 2, 0, 0, 1)]) + struct.pack(">B3I", 3, 0, 1, 2)
    (tmp_path / "t.ply").write_bytes(hdr_be + body_be)
    be = SceneLoaderManager().load(str(tmp_path / "top.pbrt"))
    assert np.array_equal(np.ctypeslib.as_array(be.desc.contents.meshes[0].P, (9,)), [0, 0, 0, 2, 0, 0, 0, 2, 0])
    with pytest.raises(SceneError, match="unknown object"):
        SceneLoaderManager().load_string('Camera "perspective" WorldBegin ObjectInstance "nope" WorldEnd', "pbrt")
    with pytest.raises(SceneError, match="cannot open"):
        SceneLoaderManager().load_string('Camera "perspective" WorldBegin Shape "plymesh" "string filename" "missing.ply" WorldEnd', "pbrt")


def test_textures_pbrt_imagemap_and_json_round_trip(tmp_path):
    """Texture "imagemap" + "texture Kd" -> BSDFColor::Bitmap on the matte / substrate diffuse slot (bsdfs/mod.rs:219-241,
    299-306, 358-374); .pfm through Bitmap::read_pfm (rows flipped back), binary .ppm as value / 255; JSON keeps every texture."""
    from rustlight_b200.host import save_pfm
    img = np.random.default_rng(3).random((3, 4, 3)).astype(np.float32)
    save_pfm(str(tmp_path / "a.pfm"), img)
    ppm = (np.arange(2 * 2 * 3, dtype=np.uint8) * 20).reshape(2, 2, 3)
    (tmp_path / "b.ppm").write_bytes(b"P6\n# comment\n2 2\n255\n" + ppm.tobytes())
    tri = 'Shape "trianglemesh" "integer indices" [0 1 2] "point P" [0 0 0 1 0 0 0 1 0] "float uv" [0 0 1 0 0 1]'
    (tmp_path / "s.pbrt").write_text(f'''Camera "perspective" WorldBegin
      Texture "ta" "spectrum" "imagemap" "string filename" "a.pfm"
      Texture "tb" "color" "imagemap" "string filename" "b.ppm"
      Material "matte" "texture Kd" "ta" {tri}
      Material "substrate" "texture Kd" "tb" "rgb Ks" [0.1 0.1 0.1] {tri}
      Material "matte" "rgb Kd" [0.3 0.3 0.3] {tri}
    WorldEnd''')
    sc = SceneLoaderManager().load(str(tmp_path / "s.pbrt"))
    d = sc.desc.contents
    assert d.ntextures == 2 and [d.meshes[i].mat.kd_texture for i in range(3)] == [1, 2, 0]
    t = d.textures[0]
    assert (t.kind, t.width, t.height) == (_abi.RL_TEX_BITMAP, 4, 3)
    assert np.array_equal(np.ctypeslib.as_array(t.pixels, (36,)).reshape(3, 4, 3), img)
    assert np.allclose(np.ctypeslib.as_array(d.textures[1].pixels, (12,)), ppm.ravel() / 255.0, rtol=1e-6)
    sc.add_checkerboard_texture((1, 0, 0), (0, 1, 0), (0.1, 0.2), (3, 4))
    sc.add_grid_texture((1, 1, 1), (0, 0, 0), 0.02, (0, 0), (2, 2))
    back_scene = SceneLoaderManager().load_string(sc.to_json(), "json")
    b = back_scene.desc.contents
    assert b.ntextures == 4 and [b.meshes[i].mat.kd_texture for i in range(3)] == [1, 2, 0]
    assert np.array_equal(np.ctypeslib.as_array(b.textures[0].pixels, (36,)), np.ctypeslib.as_array(t.pixels, (36,)))
    assert (b.textures[2].kind, list(b.textures[2].offset), list(b.textures[2].scale)) == (_abi.RL_TEX_CHECKERBOARD, pytest.approx([0.1, 0.2]), [3, 4])
    assert b.textures[3].kind == _abi.RL_TEX_GRID and b.textures[3].line_width == pytest.approx(0.02)
    with pytest.raises(SceneError, match="unknown texture"):
        SceneLoaderManager().load_string(f'Camera "perspective" WorldBegin Material "matte" "texture Kd" "zz" {tri} WorldEnd', "pbrt")


def test_pbrt_light_sources():
    """LightSource "point" / "distant" -> PointEmitter / DirectionalLight (scene_loader.rs:207-240): intensity * scale,
    direction = normalize(to - from), positions through the current transform; "infinite" stays outside the path."""
    tri = 'Shape "trianglemesh" "integer indices" [0 1 2] "point P" [0 0 0 1 0 0 0 1 0]'
    txt = f'''Camera "perspective" WorldBegin
      LightSource "point" "rgb I" [1 2 3] "rgb scale" [2 2 2] "point from" [0 1 0]
      AttributeBegin Translate 0 0 5 LightSource "point" "point from" [1 0 0] AttributeEnd
      LightSource "distant" "rgb L" [4 4 4] "point from" [0 10 0] "point to" [0 0 10]
      {tri} WorldEnd'''
    sc = SceneLoaderManager().load_string(txt, "pbrt")
    d = sc.desc.contents
    assert d.nlights == 3
    assert (d.lights[0].kind, list(d.lights[0].intensity), list(d.lights[0].v)) == (_abi.RL_LIGHT_POINT, [2, 4, 6], [0, 1, 0])
    assert list(d.lights[1].v) == [1, 0, 5] and list(d.lights[1].intensity) == [1, 1, 1]
    assert d.lights[2].kind == _abi.RL_LIGHT_DIRECTIONAL and list(d.lights[2].v) == pytest.approx([0, -2 ** -0.5, 2 ** -0.5])
    back_scene = SceneLoaderManager().load_string(sc.to_json(), "json")  # keep the owner alive: desc points into it
    back = back_scene.desc.contents
    assert back.nlights == 3 and all(bytes(back.lights[i]) == bytes(d.lights[i]) for i in range(2))
    assert list(back.lights[2].v) == pytest.approx(list(d.lights[2].v), abs=1e-7)  # re-normalised on load
    env = SceneLoaderManager().load_string('Camera "perspective" WorldBegin LightSource "infinite" "rgb L" [0.5 1 2] "rgb scale" [2 2 2] ' + tri + ' WorldEnd', "pbrt")
    de = env.desc.contents
    assert de.has_environment == 1 and list(de.environment) == [1, 2, 4]   # scene_loader.rs:241-258: L * scale
    env_back = SceneLoaderManager().load_string(env.to_json(), "json")
    assert env_back.desc.contents.has_environment == 1 and list(env_back.desc.contents.environment) == [1, 2, 4]
    with pytest.raises(SceneError, match="pfm"):  # environment textures are read from .pfm / .ppm (test_envmap.py); no EXR reader here
        SceneLoaderManager().load_string('Camera "perspective" WorldBegin LightSource "infinite" "string mapname" "sky.exr" WorldEnd', "pbrt")
    with pytest.raises(SceneError, match="scope"):
        SceneLoaderManager().load_string('Camera "perspective" WorldBegin LightSource "spot" WorldEnd', "pbrt")


def test_json_materials_round_trip():
    import json
    from rustlight_b200.host import material_glass, material_metal, material_mirror, material_substrate
    sc = load_cbox(16, 16)
    mats = [material_mirror((0.9, 0.8, 0.7)), material_metal(alpha=0.07), material_metal(microfacet="beckmann", alpha=0.2),
            material_glass(int_ior=1.5046, ext_ior=1.000277), material_substrate((0.4, 0.3, 0.2), (0.05, 0.05, 0.05), "ggx", 0.15),
            material_substrate(microfacet=None)]
    for i, m in enumerate(mats):
        sc.set_material(i, m)
    back_scene = SceneLoaderManager().load_string(sc.to_json(), "json")  # keep the owner alive: desc points into it
    back = back_scene.desc.contents
    for i, m in enumerate(mats):
        got = back.meshes[i].mat
        assert bytes(got) == bytes(m), i
    assert json.loads(sc.to_json())["meshes"][3]["material"]["type"] == "glass"


def test_pbrt_transforms_and_defaults():
    txt = '''LookAt 0 0 5  0 0 0  0 1 0
    Camera "perspective" "float fov" [30]
    Film "image" "integer xresolution" [32] "integer yresolution" [16]
    WorldBegin
      AttributeBegin
        Translate 1 2 3  Scale 2 2 2
        Shape "trianglemesh" "integer indices" [0 1 2] "point P" [0 0 0 1 0 0 0 1 0]
      AttributeEnd
      Material "phong" "rgb Kd" [0.2 0.2 0.2] "rgb Ks" [0.6 0.6 0.6] "float exponent" [10]
      Shape "trianglemesh" "integer indices" [0 1 2] "point P" [0 0 0 1 0 0 0 1 0]
    WorldEnd'''
    sc = SceneLoaderManager().load_string(txt, "pbrt")
    d = sc.desc.contents
    assert sc.size == (32, 16) and d.nmeshes == 2
    P0 = np.ctypeslib.as_array(d.meshes[0].P, (9,)).reshape(3, 3)
    assert np.array_equal(P0, [[1, 2, 3], [3, 2, 3], [1, 4, 3]])
    assert np.allclose(list(d.meshes[0].mat.kd), [0.5] * 3)          # no material: diffuse 0.5 (scene_loader.rs:132-135)
    P1 = np.ctypeslib.as_array(d.meshes[1].P, (9,)).reshape(3, 3)
    assert np.array_equal(P1, [[0, 0, 0], [1, 0, 0], [0, 1, 0]])     # AttributeEnd restored the CTM
    assert d.meshes[1].mat.kind == _abi.RL_BSDF_PHONG and d.meshes[1].mat.weight_specular == pytest.approx(0.75)
    tw = np.array(d.camera.to_world, np.float32).reshape(4, 4).T
    assert np.allclose(tw[:3, 3], [0, 0, 5], atol=1e-6) and np.allclose(tw[:3, 2], [0, 0, -1], atol=1e-6)


def test_pbrt_named_coordinate_systems():
    """pbrt-v3 CoordinateSystem / CoordSysTransform: a saved CTM comes back by name; "camera" is camera space -> world."""
    txt = '''LookAt 0 0 5  0 0 0  0 1 0
    Camera "perspective" "float fov" [30]
    WorldBegin
      Translate 1 2 3
      CoordinateSystem "shifted"
      Identity
      Shape "trianglemesh" "integer indices" [0 1 2] "point P" [0 0 0 1 0 0 0 1 0]
      CoordSysTransform "shifted"
      Shape "trianglemesh" "integer indices" [0 1 2] "point P" [0 0 0 1 0 0 0 1 0]
      CoordSysTransform "camera"
      Shape "trianglemesh" "integer indices" [0 1 2] "point P" [0 0 1 1 0 1 0 1 1]
    WorldEnd'''
    sc = SceneLoaderManager().load_string(txt, "pbrt")
    d = sc.desc.contents
    P = [np.ctypeslib.as_array(d.meshes[i].P, (9,)).reshape(3, 3) for i in range(3)]
    assert np.array_equal(P[0], [[0, 0, 0], [1, 0, 0], [0, 1, 0]])
    assert np.array_equal(P[1], [[1, 2, 3], [2, 2, 3], [1, 3, 3]])
    # camera space (pbrt looks down +z): one unit in front of the camera at (0,0,5), which looks towards -z of the world
    assert np.allclose(P[2][0], [0, 0, 4], atol=1e-5) and np.allclose(P[2][2], [0, 1, 4], atol=1e-5)
    with pytest.raises(Exception, match="unknown coordinate system"):
        SceneLoaderManager().load_string('WorldBegin CoordSysTransform "nope" WorldEnd', "pbrt")


def test_scale_image_truncates_and_keeps_matrices():
    sc = load_cbox()
    before = bytes(sc.desc.contents.camera)[8:]
    sc.scale_image(0.3)                       # CLI -s (camera.rs:73-78)
    assert sc.size == (153, 153) and bytes(sc.desc.contents.camera)[8:] == before


def test_pfm_roundtrip_and_layout(tmp_path):
    img = np.arange(2 * 3 * 3, dtype=np.float32).reshape(2, 3, 3) - 4.0
    p = str(tmp_path / "a.pfm")
    save_pfm(p, img)
    raw = open(p, "rb").read()
    assert raw.startswith(b"PF\n3 2\n-1.0\n")                           # structure.rs:550
    body = np.frombuffer(raw[len(b"PF\n3 2\n-1.0\n"):], "<f4").reshape(2, 3, 3)
    assert np.array_equal(body, np.abs(img[::-1]))                        # bottom-to-top rows of abs() (:552-558)
    assert np.array_equal(read_pfm(p), np.abs(img))


HEADER = open(os.path.join(ROOT, "include", "rl_b200.h")).read()
DECLARED = sorted(set(re.findall(r"\b(rl_[a-z_0-9]+)\s*\(", HEADER)))


def test_c_abi_exports_every_declared_symbol():
    """The CUDA library loads without a GPU and exports exactly the entry points of rl_b200.h."""
    path = os.path.join(ROOT, "rustlight_b200", "librl_b200.so")
    assert os.path.exists(path), "CUDA extension not built: python -c 'import __graft_entry__ as g; g.build()'"
    lib = C.CDLL(path)
    assert len(DECLARED) >= 14
    for name in DECLARED:
        assert hasattr(lib, name), name
    lib.rl_abi_version.restype = C.c_int
    assert lib.rl_abi_version() == int(re.search(r"#define RL_B200_ABI_VERSION (\d+)", HEADER).group(1))


def test_struct_layouts_match_the_header(tmp_path):
    """sizeof() from the C compiler vs the ctypes mirror in rustlight_b200/_abi.py."""
    import subprocess
    names = ["rl_material", "rl_mesh_desc", "rl_camera_desc", "rl_scene_desc", "rl_integrator_desc", "rl_render_opts",
             "rl_stats", "rl_bvh_info", "rl_layout_info"]
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "rl_b200.h"\nint main(void){' +
                   "".join(f'printf("%zu\\n", sizeof({n}));' for n in names) + "return 0;}")
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert sizes == [C.sizeof(getattr(_abi, n)) for n in names]


def test_no_gpu_means_loud_failure_not_fallback():
    from conftest import has_gpu
    if has_gpu():
        pytest.skip("GPU present")
    from rustlight_b200.device import Context, DeviceError
    with pytest.raises(DeviceError, match="no CUDA device"):
        Context(0)


def test_product_never_touches_the_oracle():
    """The product tree must not import, link or mention oracle/ or tests/emu."""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "rustlight_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"(from|import)\s+oracle|oracle/|liboracle|libemu|emu_binding", txt):
                    bad.append(f)
    assert not bad, bad
