"""MTSSceneLoader (scene_loader.rs:318-795) + bsdf_mts (bsdfs/mod.rs:395-612): the Mitsuba-XML subset, the reference's only route
to BSDFPhong.  mitsuba_rs is not vendored, so these tests pin the loader against the reference's own mapping code (what it does
with the parsed values) and against the Mitsuba 0.x documentation for the XML conventions (unpinned)."""
import math

import numpy as np
import pytest

import emu_binding as eb
from oracle import binding as ob
from rustlight_b200 import SceneLoaderManager, _abi
from rustlight_b200.host import SceneError, camera_create, material_phong

STREAM = dict(estimator=ob.EST_STREAM, accel_mode=ob.ACCEL_BVH)

BOX = """<?xml version="1.0" encoding="utf-8"?>
<!-- an open box of rectangles lit by an area light, Phong walls: the shape of BASELINE config 3 -->
<scene version="0.6.0">
    <default name="spp" value="16"/>
    <default name="wall" value="0.5, 0.5, 0.5"/>
    <integrator type="path"/>
    <sensor type="perspective">
        <float name="fov" value="40"/>
        <string name="fovAxis" value="x"/>
        <transform name="toWorld">
            <lookat origin="0, 1, 4.5" target="0, 1, 0" up="0, 1, 0"/>
        </transform>
        <sampler type="independent"><integer name="sampleCount" value="$spp"/></sampler>
        <film type="hdrfilm"><integer name="width" value="40"/><integer name="height" value="32"/></film>
    </sensor>
    <bsdf type="phong" id="white">
        <rgb name="diffuseReflectance" value="$wall"/>
        <spectrum name="specularReflectance" value="0.3"/>
        <float name="exponent" value="50"/>
    </bsdf>
    <bsdf type="twosided" id="red"><bsdf type="diffuse"><rgb name="reflectance" value="0.63, 0.065, 0.05"/></bsdf></bsdf>
    <texture type="checkerboard" id="checks">
        <rgb name="color0" value="0.8, 0.8, 0.8"/><rgb name="color1" value="0.1, 0.1, 0.1"/>
        <float name="uscale" value="4"/><float name="vscale" value="4"/>
    </texture>
    <shape type="rectangle" id="floor">
        <transform name="toWorld"><rotate x="1" angle="-90"/><translate y="0"/></transform>
        <bsdf type="diffuse"><ref name="reflectance" id="checks"/></bsdf>
    </shape>
    <shape type="rectangle" id="ceiling">
        <transform name="toWorld"><rotate x="1" angle="90"/><translate y="2"/></transform>
        <ref id="white"/>
    </shape>
    <shape type="rectangle" id="back">
        <transform name="toWorld"><translate y="1" z="-1"/></transform>
        <ref id="white"/>
    </shape>
    <shape type="rectangle">
        <transform name="toWorld"><rotate y="1" angle="90"/><translate x="-1" y="1"/></transform>
        <ref id="red"/>
    </shape>
    <shape type="rectangle">
        <transform name="toWorld"><rotate y="1" angle="-90"/><translate x="1" y="1"/></transform>
        <bsdf type="roughplastic"><rgb name="diffuseReflectance" value="0.14, 0.45, 0.091"/><string name="distribution" value="ggx"/><float name="alpha" value="0.2"/></bsdf>
    </shape>
    <shape type="sphere">
        <point name="center" x="-0.3" y="0.4" z="-0.2"/><float name="radius" value="0.4"/>
        <bsdf type="roughconductor"><float name="alpha" value="0.15"/><rgb name="eta" value="0.2, 0.9, 1.1"/><rgb name="k" value="3.9, 2.4, 2.1"/><float name="extEta" value="1"/></bsdf>
    </shape>
    <shape type="rectangle">
        <transform name="toWorld"><scale value="0.25"/><rotate x="1" angle="90"/><translate y="1.98"/></transform>
        <emitter type="area"><rgb name="radiance" value="17, 12, 4"/></emitter>
    </shape>
    <emitter type="point"><point name="position" x="0.5" y="1.0" z="0.5"/><rgb name="intensity" value="0.2"/></emitter>
</scene>
"""


def _load(text=BOX):
    return SceneLoaderManager().load_string(text, "xml")


def test_shapes_bsdfs_and_emitters_follow_the_reference_mapping():
    sc = _load()
    d = sc.desc.contents
    # shapes with an id come first (shapes_id then shapes_unamed, scene_loader.rs:379-383), in document order
    assert d.nmeshes == 7 and [d.meshes[i].ntris for i in range(7)] == [2, 2, 2, 2, 2, 31 * 31 * 2, 2]
    assert sc.size == (40, 32)
    m = [d.meshes[i].mat for i in range(7)]
    assert m[0].kind == _abi.RL_BSDF_DIFFUSE and m[0].kd_texture == 1 and d.ntextures == 1 and d.textures[0].kind == _abi.RL_TEX_CHECKERBOARD
    assert list(d.textures[0].scale) == [4, 4] and list(d.textures[0].color0) == pytest.approx([0.8] * 3)
    ph = material_phong((0.5, 0.5, 0.5), (0.3, 0.3, 0.3), 50.0)  # weight_specular = s_avg / (d_avg + s_avg), mod.rs:518-523
    assert m[1].kind == _abi.RL_BSDF_PHONG and bytes(m[1]) == bytes(ph) and bytes(m[2]) == bytes(ph)
    assert m[3].kind == _abi.RL_BSDF_DIFFUSE and list(m[3].kd) == pytest.approx([0.63, 0.065, 0.05])          # twosided is transparent
    assert m[4].kind == _abi.RL_BSDF_SUBSTRATE and m[4].microfacet == _abi.RL_MICROFACET_GGX and m[4].alpha == pytest.approx(0.2)
    assert list(m[4].ks) == [1, 1, 1] and list(m[4].kd) == pytest.approx([0.14, 0.45, 0.091])
    assert m[5].kind == _abi.RL_BSDF_METAL and m[5].microfacet == _abi.RL_MICROFACET_BECKMANN and list(m[5].eta) == pytest.approx([0.2, 0.9, 1.1])
    assert m[6].kind == _abi.RL_BSDF_DIFFUSE and list(m[6].kd) == pytest.approx([0.8] * 3)                    # no bsdf: BSDFDiffuse 0.8
    assert d.meshes[6].emission_kind == 1 and list(d.meshes[6].emission) == [17, 12, 4]
    assert d.nlights == 1 and d.lights[0].kind == _abi.RL_LIGHT_POINT and list(d.lights[0].v) == [0.5, 1.0, 0.5] and list(d.lights[0].intensity) == pytest.approx([0.2] * 3)
    # rectangle = (-1,-1,0) (1,-1,0) (1,1,0) (-1,1,0), indices (0,1,2) (2,3,0), normals +z, uv corners (:538-563); the floor is rotated by -90 deg about x
    P = np.ctypeslib.as_array(d.meshes[0].P, (4, 3))
    assert np.allclose(P, [[-1, 0, 1], [1, 0, 1], [1, 0, -1], [-1, 0, -1]], atol=1e-6)
    assert np.allclose(np.ctypeslib.as_array(d.meshes[0].N, (4, 3)), [[0, 1, 0]] * 4, atol=1e-6)
    assert list(np.ctypeslib.as_array(d.meshes[0].idx, (6,))) == [0, 1, 2, 2, 3, 0]
    assert np.allclose(np.ctypeslib.as_array(d.meshes[0].UV, (4, 2)), [[0, 0], [1, 0], [1, 1], [0, 1]])
    # the light: scaled to 0.5 x 0.5, facing down just below the ceiling
    L = np.ctypeslib.as_array(d.meshes[6].P, (4, 3))
    assert np.allclose(L[:, 1], 1.98, atol=1e-6) and np.allclose(np.abs(L[:, [0, 2]]), 0.25, atol=1e-6)
    assert np.allclose(np.ctypeslib.as_array(d.meshes[6].N, (4, 3)), [[0, -1, 0]] * 4, atol=1e-6)
    # sphere (:596-665): 32 x 32 vertices around the centre, unit normals, poles duplicated
    S = np.ctypeslib.as_array(d.meshes[5].P, (1024, 3))
    assert np.allclose(np.linalg.norm(S - [-0.3, 0.4, -0.2], axis=1), 0.4, atol=1e-5)
    assert np.allclose(S[0], [-0.3, 0.4, 0.2], atol=1e-6)  # theta = 0 -> +z


def test_camera_is_camera_new_with_flip(tmp_path):
    """Camera::new(img_size, Fov::X(fov), to_world, true) (scene_loader.rs:327-338); lookat = Mitsuba's camera-to-world (left, up, dir, origin)."""
    sc = _load()
    to_world = np.float32([[-1, 0, 0, 0], [0, 1, 0, 0], [0, 0, -1, 0], [0, 1, 4.5, 1]])  # columns: left = up x dir, up, dir = -z, origin
    s2c, _ = camera_create(40, 32, 40.0, to_world.ravel(), fov_axis="x", flip=True)
    cam = sc.desc.contents.camera
    assert np.allclose(np.array(cam.to_world), to_world.ravel(), atol=1e-6)
    assert np.allclose(np.array(cam.sample_to_camera), s2c, rtol=1e-6, atol=1e-7)
    # the box is in view: the centre pixel sees the back wall
    prim, tuv = ob.OracleScene(sc).primary_hits(ob.ACCEL_BVH)
    assert prim[16, 20] in (4, 5) and tuv[16, 20, 0] == pytest.approx(5.5, rel=1e-3)


def test_transform_operations_compose_in_document_order():
    xml = BOX.replace('<transform name="toWorld"><translate y="1" z="-1"/></transform>',
                      '<transform name="toWorld"><scale x="2" y="0.5"/><rotate z="1" angle="90"/><translate x="1" y="2" z="3"/><matrix value="1 0 0 0.5  0 1 0 0  0 0 1 0  0 0 0 1"/></transform>')
    sc = _load(xml)
    P = np.ctypeslib.as_array(sc.desc.contents.meshes[2].P, (4, 3))
    # (1, 1, 0): scale -> (2, 0.5, 0); rotate 90 deg about z -> (-0.5, 2, 0); translate -> (0.5, 4, 3); matrix (+0.5 in x) -> (1, 4, 3)
    assert np.allclose(P[2], [1.0, 4.0, 3.0], atol=1e-5)


@pytest.mark.parametrize("integ", [_abi.path_desc(max_depth=5), _abi.direct_desc(1, 1)], ids=["path", "direct"])
def test_xml_scene_renders_bit_exact(integ):
    """The XML route end to end: Phong walls (BSDFPhong exists only here in the reference), a checkerboard floor, a rough-plastic wall,
    a rough-conductor sphere (tree traversal: 1934 triangles), an area light and a point light -- emulator == oracle bit for bit."""
    sc = _load()
    ie, se = eb.EmuScene(sc, "sah4").render(integ, 4, seed=3)
    io, so = ob.OracleScene(sc).render(integ, 4, seed=3, cfg=ob.config(**STREAM))
    assert np.isfinite(io).all() and io.mean() > 0.02
    assert (se.segments, se.hits, se.shadow_rays) == (so.segments, so.hits, so.shadow_rays)
    assert np.array_equal(ie, io)


def test_obj_and_ply_shapes(tmp_path):
    (tmp_path / "quad.obj").write_text("v -1 0 -1\nv 1 0 -1\nv 1 0 1\nv -1 0 1\nvt 0 0\nvt 1 0\nvt 1 1\nvt 0 1\nvn 0 1 0\nf 1/1/1 2/2/1 3/3/1 4/4/1\n")
    (tmp_path / "tri.ply").write_text("ply\nformat ascii 1.0\nelement vertex 3\nproperty float x\nproperty float y\nproperty float z\nelement face 1\n"
                                      "property list uchar int vertex_indices\nend_header\n0 0 0\n1 0 0\n0 1 0\n3 0 1 2\n")
    xml = BOX.replace('<shape type="sphere">', '<shape type="obj"><string name="filename" value="quad.obj"/></shape>\n'
                      '<shape type="ply"><string name="filename" value="tri.ply"/><transform name="toWorld"><translate z="-0.5"/></transform></shape>\n<shape type="sphere">')
    (tmp_path / "s.xml").write_text(xml)
    sc = SceneLoaderManager().load(str(tmp_path / "s.xml"))  # (keep the owner alive: desc points into it)
    d = sc.desc.contents
    assert d.nmeshes == 9 and d.meshes[5].ntris == 2 and d.meshes[6].ntris == 1
    assert np.allclose(np.ctypeslib.as_array(d.meshes[5].UV, (4, 2)), [[0, 1], [1, 1], [1, 0], [0, 0]])  # flipTexCoords defaults to true: v -> 1 - v
    assert np.allclose(np.ctypeslib.as_array(d.meshes[6].P, (3, 3))[:, 2], -0.5)


def _serialized_shape(P, idx, N=None, UV=None, version=4, double=False, name="quad"):
    """One shape of a Mitsuba .serialized file, packed with struct + zlib only."""
    import struct
    import zlib
    real = "<f8" if double else "<f4"
    flags = (0x2000 if double else 0x1000) | (1 if N is not None else 0) | (2 if UV is not None else 0)
    body = struct.pack("<I", flags)
    if version == 4:
        body += name.encode() + b"\0"
    body += struct.pack("<QQ", len(P), len(idx))
    body += np.asarray(P, real).tobytes()
    if N is not None:
        body += np.asarray(N, real).tobytes()
    if UV is not None:
        body += np.asarray(UV, real).tobytes()
    body += np.asarray(idx, "<u4").tobytes()
    return struct.pack("<HH", 0x041C, version) + zlib.compress(body)


def test_serialized_shapes(tmp_path):
    """<shape type="serialized"> (scene_loader.rs:499-538): normals kept unless faceNormals, texcoords as stored, shapeIndex through the
    offset table at the end of the file, f64 records narrowed; versions 3 and 4."""
    import struct
    P = [[-1, 0, -1], [1, 0, -1], [1, 0, 1], [-1, 0, 1]]
    N = [[0, 1, 0]] * 4
    UV = [[0, 0], [1, 0], [1, 1], [0, 1]]
    for version in (3, 4):
        s0 = _serialized_shape(P, [[0, 1, 2], [0, 2, 3]], N, UV, version=version)
        s1 = _serialized_shape([[0, 0, 0], [1, 0, 0], [0, 1, 0]], [[0, 1, 2]], version=version, double=True, name="tri")
        table = struct.pack("<QQ" if version == 4 else "<II", 0, len(s0)) + struct.pack("<I", 2)
        (tmp_path / "m.serialized").write_bytes(s0 + s1 + table)
        xml = BOX.replace('<shape type="sphere">', '<shape type="serialized"><string name="filename" value="m.serialized"/></shape>\n'
                          '<shape type="serialized"><string name="filename" value="m.serialized"/><integer name="shapeIndex" value="1"/>'
                          '<boolean name="faceNormals" value="true"/><transform name="toWorld"><translate z="-0.5"/></transform></shape>\n<shape type="sphere">')
        (tmp_path / "s.xml").write_text(xml)
        sc = SceneLoaderManager().load(str(tmp_path / "s.xml"))
        d = sc.desc.contents
        assert d.nmeshes == 9 and d.meshes[5].ntris == 2 and d.meshes[6].ntris == 1
        assert np.array_equal(np.ctypeslib.as_array(d.meshes[5].P, (4, 3)), np.array(P, np.float32))
        assert np.array_equal(np.ctypeslib.as_array(d.meshes[5].N, (4, 3)), np.array(N, np.float32))
        assert np.array_equal(np.ctypeslib.as_array(d.meshes[5].UV, (4, 2)), np.array(UV, np.float32))  # no flip
        assert not d.meshes[6].N and not d.meshes[6].UV
        assert np.allclose(np.ctypeslib.as_array(d.meshes[6].P, (3, 3))[:, 2], -0.5)
    bad = s0[:-7]
    (tmp_path / "m.serialized").write_bytes(bad)
    with pytest.raises(SceneError, match="serialized"):
        SceneLoaderManager().load(str(tmp_path / "s.xml"))
    (tmp_path / "m.serialized").write_bytes(b"\x1c\x05" + s0[2:])
    with pytest.raises(SceneError, match="magic"):
        SceneLoaderManager().load(str(tmp_path / "s.xml"))


def test_rejected_inputs():
    with pytest.raises(SceneError, match="sensor"):
        _load(BOX.replace("<sensor", "<zsensor").replace("</sensor>", "</zsensor>"))
    with pytest.raises(SceneError, match="Fov axis"):
        _load(BOX.replace('value="x"', 'value="diagonal"'))
    with pytest.raises(SceneError, match="todo"):
        _load(BOX.replace('<emitter type="point">', '<emitter type="pointnormal">'))
    with pytest.raises(SceneError, match="media"):
        _load(BOX.replace("<integrator", '<medium type="homogeneous" id="fog"/><integrator'))
    with pytest.raises(SceneError, match="refers to nothing"):
        _load(BOX.replace('<ref id="red"/>', '<ref id="blue"/>'))
    with pytest.raises(SceneError, match="alpha_u"):
        _load(BOX.replace('<float name="alpha" value="0.15"/>', '<float name="alphaU" value="0.1"/><float name="alphaV" value="0.2"/>'))
