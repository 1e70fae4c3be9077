"""Python face of the C++ host layer (librl_host.so): scene loading and camera construction.

Mirrors the reference names: SceneLoaderManager.load (src/scene_loader.rs:28-45),
Scene.nb_samples/output_img (src/scene.rs:33-44), Camera.scale_image (src/camera.rs:73-78).
No rendering happens here.
"""
import ctypes as C
import os

import numpy as np

from . import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "librl_host.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(path)
        L.rlh_load_scene.restype = C.c_void_p
        L.rlh_load_scene.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_size_t]
        L.rlh_load_scene_string.restype = C.c_void_p
        L.rlh_load_scene_string.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_char_p, C.c_size_t]
        L.rlh_scene_free.argtypes = [C.c_void_p]
        L.rlh_scene_desc.restype = C.POINTER(_abi.rl_scene_desc)
        L.rlh_scene_desc.argtypes = [C.c_void_p]
        L.rlh_scene_nb_meshes.restype = C.c_uint32
        L.rlh_scene_nb_meshes.argtypes = [C.c_void_p]
        L.rlh_scene_nb_triangles.restype = C.c_uint64
        L.rlh_scene_nb_triangles.argtypes = [C.c_void_p]
        L.rlh_scene_scale_image.argtypes = [C.c_void_p, C.c_float]
        L.rlh_scene_set_resolution.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_char_p, C.c_size_t]
        f3 = C.c_float * 3
        L.rlh_material_metal.argtypes = [f3, f3, f3, C.c_uint32, C.c_float, C.POINTER(_abi.rl_material)]
        L.rlh_material_glass.argtypes = [f3, f3, C.c_float, C.c_float, C.POINTER(_abi.rl_material)]
        L.rlh_material_substrate.argtypes = [f3, f3, C.c_uint32, C.c_float, C.POINTER(_abi.rl_material)]
        L.rlh_remap_roughness.restype = C.c_float
        L.rlh_remap_roughness.argtypes = [C.c_float, C.c_int]
        L.rlh_scene_add_texture.restype = C.c_uint32
        L.rlh_scene_add_texture.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.rlh_scene_add_texture_file.restype = C.c_uint32
        L.rlh_scene_add_texture_file.argtypes = [C.c_void_p, C.c_char_p]
        L.rlh_scene_set_environment.argtypes = [C.c_void_p, C.c_float * 3]
        L.rlh_scene_set_environment_texture.argtypes = [C.c_void_p, C.c_uint32]
        L.rlh_scene_override_lights.argtypes = [C.c_void_p, C.c_uint32]
        L.rlh_scene_set_ats.argtypes = [C.c_void_p, C.c_int]
        L.rlh_scene_set_ats.restype = None
        L.rlh_scene_add_light.argtypes = [C.c_void_p, C.c_uint32, C.c_float * 3, C.c_float * 3]
        L.rlh_scene_set_material.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(_abi.rl_material)]
        L.rlh_scene_set_material_blend.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(_abi.rl_material), C.POINTER(_abi.rl_material), C.c_float]
        L.rlh_material_phong.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_float,
                                         C.POINTER(_abi.rl_material)]
        L.rlh_scene_to_json.restype = C.c_size_t
        L.rlh_scene_to_json.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
        L.rlh_camera_create.argtypes = [C.c_uint32, C.c_uint32, C.c_int, C.c_float, C.POINTER(C.c_float),
                                        C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.rlh_save_pfm.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_float)]
        L.rlh_read_pfm.argtypes = [C.c_char_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                   C.POINTER(C.c_float), C.c_size_t]
        L.rlh_save_image.argtypes = L.rlh_save_pfm.argtypes
        L.rlh_read_image.argtypes = L.rlh_read_pfm.argtypes
        _lib = L
    return _lib


class SceneError(RuntimeError):
    pass


class Scene:
    """src/scene.rs:16-30 -- owns the C++ Scene; `desc` is the flat rl_scene_desc view of it."""

    def __init__(self, handle):
        self._h = handle
        self.nb_samples = 1
        self.output_img_path = "out.pfm"

    def __del__(self):
        if getattr(self, "_h", None):
            lib().rlh_scene_free(self._h)
            self._h = None

    @property
    def desc(self):
        return lib().rlh_scene_desc(self._h)

    @property
    def size(self):
        d = self.desc.contents.camera
        return int(d.width), int(d.height)

    @property
    def nb_meshes(self):
        return int(lib().rlh_scene_nb_meshes(self._h))

    @property
    def nb_triangles(self):
        return int(lib().rlh_scene_nb_triangles(self._h))

    def scale_image(self, s):
        """CLI `-s`: Camera::scale_image (truncating, matrices untouched)."""
        lib().rlh_scene_scale_image(self._h, float(s))
        return self

    def set_resolution(self, w, h):
        """Edit the Film resolution and rebuild the camera (NOT the `-s` flag)."""
        err = C.create_string_buffer(512)
        if lib().rlh_scene_set_resolution(self._h, int(w), int(h), err, 512) != 0:
            raise SceneError(err.value.decode())
        return self

    def set_material(self, mesh, material):
        if lib().rlh_scene_set_material(self._h, int(mesh), C.byref(material)) != 0:
            raise SceneError("bad mesh index")
        return self

    def set_material_blend(self, mesh, a, b, weight):
        """BSDFBlend { bsdf1: a, bsdf2: b, weight } (bsdfs/blend.rs): both parts rough (blend.rs:17), constant colours."""
        if lib().rlh_scene_set_material_blend(self._h, int(mesh), C.byref(a), C.byref(b), float(weight)) != 0:
            raise SceneError("blend: bad mesh index, a smooth / textured / nested part, or weight outside [0, 1]")
        return self

    def add_bitmap_texture(self, rgb):
        """BSDFColor::Bitmap from an (h, w, 3) float array (Bitmap.colors order: index y*w + x).  Returns the id for
        rl_material.kd_texture."""
        a = np.ascontiguousarray(rgb, dtype=np.float32)
        h, w, _ = a.shape
        t = lib().rlh_scene_add_texture(self._h, _abi.RL_TEX_BITMAP, w, h, a.ctypes.data_as(C.POINTER(C.c_float)), None)
        if not t:
            raise SceneError("bad bitmap texture")
        return t

    def add_texture_file(self, filename):
        t = lib().rlh_scene_add_texture_file(self._h, os.fsencode(filename))
        if not t:
            raise SceneError(f"cannot read texture {filename}")
        return t

    def add_checkerboard_texture(self, color0, color1, offset=(0, 0), scale=(1, 1)):
        p = (C.c_float * 11)(*color0, *color1, *offset, *scale, 0.0)
        return lib().rlh_scene_add_texture(self._h, _abi.RL_TEX_CHECKERBOARD, 0, 0, None, p)

    def add_grid_texture(self, color0, color1, line_width=0.01, offset=(0, 0), scale=(1, 1)):
        p = (C.c_float * 11)(*color0, *color1, *offset, *scale, float(line_width))
        return lib().rlh_scene_add_texture(self._h, _abi.RL_TEX_GRID, 0, 0, None, p)

    def set_environment(self, rgb):
        """Constant EnvironmentLight (emitter.rs:428-568, pbrt LightSource "infinite" "rgb L")."""
        lib().rlh_scene_set_environment(self._h, (C.c_float * 3)(*rgb))
        return self

    def set_ats(self, on=True):
        """Scene::build_emitters(build_ats) (`-x ats`): sample lights through the light tree LightSamplerATS (emitter.rs:1130-1400)."""
        lib().rlh_scene_set_ats(self._h, 1 if on else 0)
        return self

    def override_lights_hsv(self):
        """`-x hvs-light` (examples/cli.rs:410-421): every mesh light becomes EmissionType::HSV { scale = luminance of its colour }."""
        lib().rlh_scene_override_lights(self._h, 0)
        return self

    def override_lights_texture(self, tex_id):
        """`-x texture-light` (examples/cli.rs:421-427): EmissionType::Texture { scale, img } with a bitmap texture id from add_bitmap_texture."""
        if lib().rlh_scene_override_lights(self._h, int(tex_id)) != 0:
            raise SceneError("texture lights: not a bitmap texture id")
        return self

    def set_environment_texture(self, tex_id):
        """EnvironmentLightColor::new_texture(image) (emitter.rs:341-353; pbrt LightSource "infinite" "string mapname"): a lat-long
        bitmap texture id from add_bitmap_texture."""
        if lib().rlh_scene_set_environment_texture(self._h, int(tex_id)) != 0:
            raise SceneError("environment texture: not a bitmap texture id")
        return self

    def add_point_light(self, intensity, position):
        """PointEmitter (emitter.rs:186-250), appended to Scene.emitters."""
        if lib().rlh_scene_add_light(self._h, _abi.RL_LIGHT_POINT, (C.c_float * 3)(*intensity), (C.c_float * 3)(*position)) != 0:
            raise SceneError("bad light")
        return self

    def add_directional_light(self, intensity, direction):
        """DirectionalLight (emitter.rs:96-190); `direction` points from the light into the scene and is normalised."""
        if lib().rlh_scene_add_light(self._h, _abi.RL_LIGHT_DIRECTIONAL, (C.c_float * 3)(*intensity), (C.c_float * 3)(*direction)) != 0:
            raise SceneError("bad light")
        return self

    def mesh_is_light(self, mesh):
        return bool(self.desc.contents.meshes[mesh].emission_kind)

    def to_json(self):
        n = lib().rlh_scene_to_json(self._h, None, 0)
        buf = C.create_string_buffer(n)
        lib().rlh_scene_to_json(self._h, buf, n)
        return buf.value.decode()


class SceneLoaderManager:
    """src/scene_loader.rs:21-58: extension -> loader ("pbrt", plus the new "json")."""

    def load(self, filename, use_shading_normal=True):
        err = C.create_string_buffer(1024)
        h = lib().rlh_load_scene(os.fsencode(filename), 1 if use_shading_normal else 0, err, 1024)
        if not h:
            raise SceneError(err.value.decode())
        return Scene(h)

    def load_string(self, text, fmt, use_shading_normal=True):
        err = C.create_string_buffer(1024)
        h = lib().rlh_load_scene_string(text.encode(), fmt.encode(), 1 if use_shading_normal else 0, err, 1024)
        if not h:
            raise SceneError(err.value.decode())
        return Scene(h)


def material_phong(kd, ks, exponent):
    m = _abi.rl_material()
    a = (C.c_float * 3)(*kd)
    b = (C.c_float * 3)(*ks)
    if lib().rlh_material_phong(a, b, float(exponent), C.byref(m)) != 0:
        raise SceneError("Phong: kd and ks are both black")
    return m


_MICROFACET = {None: _abi.RL_MICROFACET_NONE, "none": _abi.RL_MICROFACET_NONE, "ggx": _abi.RL_MICROFACET_GGX,
               "beckmann": _abi.RL_MICROFACET_BECKMANN}


def _f3(v):
    return (C.c_float * 3)(*v)


def material_metal(specular=(1, 1, 1), eta=(0.2004, 0.9240, 1.1022), k=(3.9129, 2.4528, 2.1421), microfacet="ggx", alpha=0.1):
    """BSDFMetal (bsdfs/metal.rs); microfacet=None is the pure specular lobe."""
    m = _abi.rl_material()
    if lib().rlh_material_metal(_f3(specular), _f3(eta), _f3(k), _MICROFACET[microfacet], float(alpha), C.byref(m)) != 0:
        raise SceneError("metal: bad parameters")
    return m


def material_mirror(kr=(0.9, 0.9, 0.9)):
    """pbrt "mirror" = BSDFMetal{specular: Kr, eta: 1, k: 0, distribution: None} (bsdfs/mod.rs:349-357)."""
    return material_metal(kr, (1, 1, 1), (0, 0, 0), None, 0.0)


def material_glass(reflectance=(1, 1, 1), transmittance=(1, 1, 1), int_ior=1.5046, ext_ior=1.000277):
    """BSDFGlass.eta(int_ior, ext_ior) (bsdfs/glass.rs:43-72)."""
    m = _abi.rl_material()
    if lib().rlh_material_glass(_f3(reflectance), _f3(transmittance), float(int_ior), float(ext_ior), C.byref(m)) != 0:
        raise SceneError("glass: eta must not be 0")
    return m


def material_substrate(diffuse=(0.5, 0.5, 0.5), specular=(0.5, 0.5, 0.5), microfacet="ggx", alpha=0.1):
    """BSDFSubstrate (bsdfs/substrate.rs)."""
    m = _abi.rl_material()
    if lib().rlh_material_substrate(_f3(diffuse), _f3(specular), _MICROFACET[microfacet], float(alpha), C.byref(m)) != 0:
        raise SceneError("substrate: bad parameters")
    return m


def remap_roughness(v, remap=True):
    return lib().rlh_remap_roughness(float(v), 1 if remap else 0)


def material_diffuse(kd=(0.0, 0.0, 0.0), kd_texture=0):
    m = _abi.rl_material()
    m.kind = _abi.RL_BSDF_DIFFUSE
    m.kd[:] = kd
    m.kd_texture = kd_texture
    return m


def camera_create(w, h, fov_deg, to_world, fov_axis="y", flip=False):
    """Camera::new -> (sample_to_camera, camera_to_sample) as 16-float column-major arrays."""
    tw = (C.c_float * 16)(*np.asarray(to_world, dtype=np.float32).ravel())
    s2c = (C.c_float * 16)()
    c2s = (C.c_float * 16)()
    if lib().rlh_camera_create(w, h, 1 if fov_axis == "x" else 0, float(fov_deg), tw, 1 if flip else 0, s2c, c2s) != 0:
        raise SceneError("Camera::new failed")
    return np.array(s2c, dtype=np.float32), np.array(c2s, dtype=np.float32)


def save_pfm(path, img):
    img = np.ascontiguousarray(img, dtype=np.float32)
    h, w, _ = img.shape
    if lib().rlh_save_pfm(os.fsencode(path), w, h, img.ctypes.data_as(C.POINTER(C.c_float))) != 0:
        raise SceneError(f"cannot write {path}")


def save_image(path, img):
    """Bitmap::save (structure.rs:528-545): by extension, .pfm or .png (gamma 2.2, 8 bit)."""
    img = np.ascontiguousarray(img, dtype=np.float32)
    h, w = img.shape[:2]
    if lib().rlh_save_image(os.fsencode(path), w, h, img.ctypes.data_as(C.POINTER(C.c_float))) != 0:
        raise SceneError(f"cannot write {path}")


def read_image(path):
    """Bitmap::read (structure.rs:670-683): by extension, .pfm or .png (8-bit RGB / 255)."""
    w, h = C.c_uint32(), C.c_uint32()
    if lib().rlh_read_image(os.fsencode(path), C.byref(w), C.byref(h), None, 0) != 0:
        raise SceneError(f"cannot read {path}")
    img = np.zeros((h.value, w.value, 3), dtype=np.float32)
    lib().rlh_read_image(os.fsencode(path), C.byref(w), C.byref(h), img.ctypes.data_as(C.POINTER(C.c_float)), img.size)
    return img


def read_pfm(path):
    w, h = C.c_uint32(), C.c_uint32()
    if lib().rlh_read_pfm(os.fsencode(path), C.byref(w), C.byref(h), None, 0) != 0:
        raise SceneError(f"cannot read {path}")
    img = np.zeros((h.value, w.value, 3), dtype=np.float32)
    lib().rlh_read_pfm(os.fsencode(path), C.byref(w), C.byref(h), img.ctypes.data_as(C.POINTER(C.c_float)), img.size)
    return img
