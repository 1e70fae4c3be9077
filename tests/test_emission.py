"""EmissionType::HSV / EmissionType::Texture (geometry.rs:99-104, Mesh::emit :184-206): what `-x hvs-light` / `-x texture-light` turn every
mesh light into (examples/cli.rs:410-429).  The emission then depends on the uv of the hit (arrival emission: vertex.rs:73, direct.rs:45,
179) and on the uv of the sampled point (light sampling: emitter.rs:637, 675 with sample_tri's NORMALIZED uv, geometry.rs:316-325), while
Mesh::flux keeps Color::value(scale) (emitter.rs:591-599).

(1) host: the override follows the CLI; (2) the oracle against the model; (3) the device arithmetic (tests/emu) == the oracle's
independently typed restatement, bit for bit; (4) what is refused.  GPU: test_gpu.py::test_varying_emission_bit_exact."""
import json
import os
import subprocess

import numpy as np
import pytest

import emu_binding as eb
from conftest import ROOT, load_cbox, rel_l2
from oracle import binding as ob
from rustlight_b200 import SceneLoaderManager, _abi
from rustlight_b200.host import SceneError

STREAM = dict(estimator=ob.EST_STREAM, accel_mode=ob.ACCEL_BVH)


def butterfly(w=8, h=6, seed=5):
    rng = np.random.default_rng(seed)
    return rng.uniform(0.0, 1.0, size=(h, w, 3)).astype(np.float32)


def hsv_box(w=32, h=32):
    return load_cbox(w, h).override_lights_hsv()


def texture_box(w=32, h=32):
    sc = load_cbox(w, h)
    return sc.override_lights_texture(sc.add_bitmap_texture(butterfly()))


# ---- (1) host -----------------------------------------------------------------------------------------------------------------------------
def test_override_follows_the_cli():
    sc = hsv_box()
    d = sc.desc.contents
    lum = np.float32(17.0) * np.float32(0.212671) + np.float32(12.0) * np.float32(0.715160) + np.float32(4.0) * np.float32(0.072169)  # Color::luminance
    assert [d.meshes[i].emission_kind for i in range(8)] == [0] * 7 + [_abi.RL_EMISSION_HSV]
    assert d.meshes[7].emission[0] == lum and d.meshes[7].emission_texture == 0
    sc.override_lights_hsv()  # `_ => 1.0`: a light that is no longer EmissionType::Color gets scale 1
    assert sc.desc.contents.meshes[7].emission[0] == 1.0
    sc = texture_box()
    d = sc.desc.contents
    assert d.meshes[7].emission_kind == _abi.RL_EMISSION_TEXTURE and d.meshes[7].emission[0] == lum and d.meshes[7].emission_texture == 1
    with pytest.raises(SceneError):
        load_cbox(8, 8).override_lights_texture(3)
    # JSON keeps both kinds
    for s in (hsv_box(), texture_box()):
        again = SceneLoaderManager().load_string(s.to_json(), "json")
        a, b = s.desc.contents.meshes[7], again.desc.contents.meshes[7]
        assert (a.emission_kind, a.emission[0], a.emission_texture) == (b.emission_kind, b.emission[0], b.emission_texture)


# ---- (2) the oracle against the model --------------------------------------------------------------------------------------------------------
def test_emission_seen_by_the_camera_is_mesh_emit():
    """A camera looking straight at the lamp (max_depth 1: sensor edge only) sees Mesh::emit(uv): the red-to-green ramp along u for
    HSV, the texels for Texture."""
    scene = {"camera": {"width": 16, "height": 16, "fov": 40.0, "to_world": [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, -1, 0, 0, 0, 3, 1]},
             "meshes": [{"material": {"type": "diffuse", "kd": [0, 0, 0]}, "emission": [2, 2, 2], "indices": [0, 1, 2, 0, 2, 3],
                         "P": [-1, -1, 0, 1, -1, 0, 1, 1, 0, -1, 1, 0], "uv": [0, 0, 1, 0, 1, 1, 0, 1]}]}
    sc = SceneLoaderManager().load_string(json.dumps(scene), "json").override_lights_hsv()
    scale = sc.desc.contents.meshes[0].emission[0]
    img, _ = ob.OracleScene(sc).render(_abi.path_desc(max_depth=2), 64, seed=1, cfg=ob.config(**STREAM))
    inner = img[2:14, 2:14]  # the lamp covers the middle of the frame (the outermost pixels see its rim)
    # r + g = scale on every pixel that sees the lamp, r follows u (screen x, possibly mirrored), b = 0
    assert np.allclose(inner[..., 0] + inner[..., 1], scale, rtol=1e-5) and not img[..., 2].any()
    cols = inner[..., 0].mean(axis=0)
    assert np.all(np.diff(cols) > 0) or np.all(np.diff(cols) < 0)
    assert np.ptp(inner[..., 0].mean(axis=1)) < 0.02 * scale  # no dependence on v
    tex = butterfly(4, 4)
    sc = SceneLoaderManager().load_string(json.dumps(scene), "json")
    sc.override_lights_texture(sc.add_bitmap_texture(tex))
    img, _ = ob.OracleScene(sc).render(_abi.path_desc(max_depth=2), 1, seed=1, cfg=ob.config(**STREAM))
    seen = {tuple(np.round(p / scale, 5)) for p in img[2:14, 2:14].reshape(-1, 3)}  # one sample per pixel: every value is one texel times scale
    texels = {tuple(np.round(p, 5)) for p in tex.reshape(-1, 3)}
    assert len(seen) >= 8 and seen <= texels


def test_flux_keeps_the_scale():
    """Mesh::flux = area * Color::value(scale) * PI for both kinds (emitter.rs:591-599): with two lamps the choice between them follows the
    scales, not what the texture holds."""
    def two_lamps(kind):
        lamp = {"material": {"type": "diffuse", "kd": [0, 0, 0]}, "indices": [0, 1, 2, 0, 2, 3], "uv": [0, 0, 1, 0, 1, 1, 0, 1]}
        scene = {"camera": {"width": 4, "height": 4, "fov": 40.0, "to_world": [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, -1, 0, 0, 0, 3, 1]},
                 "textures": {"tex1": {"type": "bitmap", "width": 1, "height": 1, "pixels": [0.01, 0.01, 0.01]}},
                 "meshes": [dict(lamp, P=[-1, -1, 0, 0, -1, 0, 0, 0, 0, -1, 0, 0], **{kind: [3.0] if kind == "emission_hsv" else {"texture": "tex1", "scale": [3.0]}}),
                            dict(lamp, P=[0, 0, 0, 1, 0, 0, 1, 1, 0, 0, 1, 0], **{kind: [1.0] if kind == "emission_hsv" else {"texture": "tex1", "scale": [1.0]}})]}
        return SceneLoaderManager().load_string(json.dumps(scene), "json")
    for kind in ("emission_hsv", "emission_texture"):
        osc = ob.OracleScene(two_lamps(kind))
        picks = [osc.sample_light(np.float32([0.0, 0.0, 2.0]), (k + 0.5) / 400, 0.3, 0.4, 0.6)["mesh"] for k in range(400)]
        assert abs(np.mean(np.array(picks) == 0) - 0.75) < 0.01


# ---- (3) emulator == oracle -------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("make", [hsv_box, texture_box])
@pytest.mark.parametrize("integ", [_abi.path_desc(), _abi.path_desc(strategy=_abi.RL_STRATEGY_EMITTER), _abi.path_desc(strategy=_abi.RL_STRATEGY_BSDF, max_depth=4),
                                   _abi.direct_desc(2, 2), _abi.direct_desc(0, 1)], ids=["path", "path-emitter", "path-bsdf", "direct22", "direct01"])
def test_renders_emulator_equals_oracle(make, integ):
    sc = make(24, 24)
    ie, se = eb.EmuScene(sc).render(integ, 6, seed=3)
    io, so = ob.OracleScene(sc).render(integ, 6, seed=3, cfg=ob.config(**STREAM))
    assert np.isfinite(io).all() and io.mean() > 0.01
    assert (se.segments, se.hits, se.shadow_rays) == (so.segments, so.hits, so.shadow_rays)
    assert np.array_equal(ie, io)


@pytest.mark.parametrize("make", [hsv_box, texture_box])
def test_stream_estimator_equals_graph(make):
    osc = ob.OracleScene(make(20, 20))
    a, sa = osc.render(_abi.path_desc(max_depth=5), 8, seed=2, cfg=ob.config(**STREAM))
    b, sb = osc.render(_abi.path_desc(max_depth=5), 8, seed=2, cfg=ob.config(estimator=ob.EST_GRAPH, accel_mode=ob.ACCEL_BVH))
    assert (sa.segments, sa.shadow_rays) == (sb.segments, sb.shadow_rays) and rel_l2(a, b) < 1e-6


def test_strategies_agree_and_the_light_tree_takes_it():
    """Light sampling (sampled, normalized uv) and BSDF sampling (uv of the hit) estimate the same image only where both see the same
    emission; the reference's normalized sample uv makes them differ by design -- what must hold is that each strategy is self-consistent
    between emulator and oracle (above) and that the tree (centroid uv for the proxies) renders too."""
    sc = hsv_box(20, 20).set_ats(True)
    ie, se = eb.EmuScene(sc).render(_abi.path_desc(), 4, seed=7)
    io, so = ob.OracleScene(sc).render(_abi.path_desc(), 4, seed=7, cfg=ob.config(**STREAM))
    assert se.segments == so.segments and np.array_equal(ie, io) and io.mean() > 0.01


# ---- (4) refused ------------------------------------------------------------------------------------------------------------------------------
def test_lights_without_uv_are_refused():
    scene = {"camera": {"width": 4, "height": 4, "fov": 40.0, "to_world": [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, -1, 0, 0, 0, 3, 1]},
             "meshes": [{"material": {"type": "diffuse", "kd": [0, 0, 0]}, "emission": [2, 2, 2], "indices": [0, 1, 2], "P": [-1, -1, 0, 1, -1, 0, 1, 1, 0]}]}
    sc = SceneLoaderManager().load_string(json.dumps(scene), "json").override_lights_hsv()
    with pytest.raises(Exception, match="uv"):
        ob.OracleScene(sc)
    with pytest.raises(Exception, match="uv"):
        eb.EmuScene(sc)


def test_cli_options(tmp_path):
    cli = os.path.join(ROOT, "rustlight_b200", "rustlight-b200")
    cbox = os.path.join(ROOT, "data", "cbox.pbrt")
    r = subprocess.run([cli, "-x", "texture-light", "-o", str(tmp_path / "o.pfm"), cbox, "path"], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode != 0 and "butterfly" in r.stderr  # no butterfly.png in the working directory
    r = subprocess.run([cli, "-x", "hvs-light", "-o", str(tmp_path / "o.pfm"), cbox, "path"], capture_output=True, text=True, cwd=tmp_path)
    assert "unknown" not in r.stderr and "outside" not in r.stderr  # accepted; without a GPU the run ends at rl_create
