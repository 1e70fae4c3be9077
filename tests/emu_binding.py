"""ctypes binding of tests/emu (serial CPU emulation of the device arithmetic; test-only)."""
import ctypes as C
import os
import subprocess

import numpy as np

from rustlight_b200 import _abi

_HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu")
_SO = os.path.join(_HERE, "_build", "libemu.so")
FP = C.POINTER(C.c_float)
_lib = None


class emu_stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("samples", "segments", "shadow_rays", "shadow_traced", "shadow_visible",
                                          "hits", "max_depth_seen")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


def lib():
    global _lib
    if _lib is None:
        subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        L = C.CDLL(_SO)
        L.emu_scene_create.restype = C.c_void_p
        L.emu_scene_create.argtypes = [C.POINTER(_abi.rl_scene_desc), C.c_char_p, C.c_size_t]
        L.emu_scene_destroy.argtypes = [C.c_void_p]
        L.emu_bvh_max_depth.restype = C.c_uint32
        L.emu_bvh_max_depth.argtypes = [C.c_void_p]
        L.emu_bvh_validate.argtypes = [C.c_void_p]
        L.emu_flat_info.restype = C.c_uint32
        L.emu_flat_info.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), FP]
        L.emu_trace.argtypes = [C.c_void_p, C.c_size_t, FP, FP, C.POINTER(C.c_uint32), FP]
        L.emu_visible.argtypes = [C.c_void_p, C.c_size_t, FP, FP, C.POINTER(C.c_uint8)]
        L.emu_primary_hits.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), FP]
        L.emu_render.argtypes = [C.c_void_p, C.POINTER(_abi.rl_integrator_desc), C.c_uint32, C.c_uint64, C.c_uint32,
                                 C.c_uint32, FP, C.POINTER(emu_stats)]
        L.emu_bsdf_sample.argtypes = [C.POINTER(_abi.rl_material), FP, C.c_float, C.c_float, FP, FP, FP]
        L.emu_bsdf_pdf.restype = C.c_float
        L.emu_bsdf_pdf.argtypes = [C.POINTER(_abi.rl_material), FP, FP]
        L.emu_bsdf_eval.argtypes = [C.POINTER(_abi.rl_material), FP, FP, FP]
        L.emu_bsdf_flags.argtypes = [C.POINTER(_abi.rl_material)]
        L.emu_ats_sample.argtypes = [C.c_void_p, C.c_float, FP, FP, C.POINTER(C.c_uint32), FP]
        L.emu_ats_pdf.argtypes = [C.c_void_p, C.c_uint32, FP, FP, C.c_int, FP]
        L.emu_spec_atan2.restype = C.c_float
        L.emu_spec_atan2.argtypes = [C.c_float, C.c_float]
        L.emu_spec_acos.restype = C.c_float
        L.emu_spec_acos.argtypes = [C.c_float]
        L.emu_env_eval_pdf.argtypes = [C.c_void_p, FP, FP, FP]
        L.emu_env_sample.argtypes = [C.c_void_p, C.c_float, C.c_float, FP, FP, FP]
        _lib = L
    return _lib


def bsdf_sample_ex(mat, wi, s0, s1):
    wi = np.ascontiguousarray(wi, np.float32)
    w, d = np.zeros(3, np.float32), np.zeros(3, np.float32)
    pdf = C.c_float()
    rc = lib().emu_bsdf_sample(C.byref(mat), _f(wi), s0, s1, _f(w), _f(d), C.byref(pdf))
    return (rc != 0, w, d, pdf.value, rc == 2)


def bsdf_pdf(mat, wi, wo):
    wi, wo = np.ascontiguousarray(wi, np.float32), np.ascontiguousarray(wo, np.float32)
    return lib().emu_bsdf_pdf(C.byref(mat), _f(wi), _f(wo))


def bsdf_eval(mat, wi, wo):
    wi, wo = np.ascontiguousarray(wi, np.float32), np.ascontiguousarray(wo, np.float32)
    out = np.zeros(3, np.float32)
    lib().emu_bsdf_eval(C.byref(mat), _f(wi), _f(wo), _f(out))
    return out


def bsdf_flags(mat):
    f = lib().emu_bsdf_flags(C.byref(mat))
    return dict(twosided=bool(f & 1), smooth=bool(f & 2))


def _f(a):
    return a.ctypes.data_as(FP)


class EmuScene:
    def __init__(self, scene, accel=None):
        """accel: None/"flat" (group table when the scene is small enough, the device default for incoherent rays),
        "leaf" (whole scene as one leaf of single-triangle records), "tree" (LBVH with small leaves)."""
        self._scene = scene
        err = C.create_string_buffer(512)
        old = os.environ.get("RL_EMU_ACCEL")
        if accel:
            os.environ["RL_EMU_ACCEL"] = accel
        try:
            self._h = lib().emu_scene_create(scene.desc, err, 512)
        finally:
            if accel:
                if old is None:
                    del os.environ["RL_EMU_ACCEL"]
                else:
                    os.environ["RL_EMU_ACCEL"] = old
        if not self._h:
            raise RuntimeError("emu: " + err.value.decode())
        self.width, self.height = scene.size

    def __del__(self):
        if getattr(self, "_h", None):
            lib().emu_scene_destroy(self._h)
            self._h = None

    def bvh_max_depth(self):
        return lib().emu_bvh_max_depth(self._h)

    def flat_info(self):
        pairs, singles, delta = C.c_uint32(), C.c_uint32(), C.c_float()
        g = lib().emu_flat_info(self._h, C.byref(pairs), C.byref(singles), C.byref(delta))
        return dict(groups=g, pairs=pairs.value, singles=singles.value, delta=delta.value)

    def bvh_validate(self):
        return lib().emu_bvh_validate(self._h)

    def ats_sample(self, r, p, n):
        prim, pdf = C.c_uint32(), C.c_float()
        if lib().emu_ats_sample(self._h, float(r), _f(np.ascontiguousarray(p, np.float32)), _f(np.ascontiguousarray(n, np.float32)), C.byref(prim), C.byref(pdf)) != 0:
            raise ValueError("no light tree")
        return prim.value, pdf.value

    def ats_pdf(self, prim, p, n=None):
        pdf = C.c_float()
        nn = np.zeros(3, np.float32) if n is None else np.ascontiguousarray(n, np.float32)
        if lib().emu_ats_pdf(self._h, int(prim), _f(np.ascontiguousarray(p, np.float32)), _f(nn), 0 if n is None else 1, C.byref(pdf)) != 0:
            raise ValueError("no light tree / not a light")
        return pdf.value

    def env_eval_pdf(self, d):
        rgb, pdf = np.zeros(3, np.float32), C.c_float()
        if lib().emu_env_eval_pdf(self._h, _f(np.ascontiguousarray(d, np.float32)), _f(rgb), C.byref(pdf)) != 0:
            raise ValueError("no environment texture")
        return rgb, pdf.value

    def env_sample(self, u0, u1):
        d, rgb, pdf = np.zeros(3, np.float32), np.zeros(3, np.float32), C.c_float()
        if lib().emu_env_sample(self._h, float(u0), float(u1), _f(d), _f(rgb), C.byref(pdf)) != 0:
            raise ValueError("no environment texture")
        return d, rgb, pdf.value

    def trace(self, o, d):
        o = np.ascontiguousarray(o, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(d, np.float32).reshape(-1, 3)
        prim = np.zeros(o.shape[0], np.uint32)
        tuv = np.zeros((o.shape[0], 3), np.float32)
        lib().emu_trace(self._h, o.shape[0], _f(o), _f(d), prim.ctypes.data_as(C.POINTER(C.c_uint32)), _f(tuv))
        return prim, tuv

    def visible(self, p0, p1):
        p0 = np.ascontiguousarray(p0, np.float32).reshape(-1, 3)
        p1 = np.ascontiguousarray(p1, np.float32).reshape(-1, 3)
        out = np.zeros(p0.shape[0], np.uint8)
        lib().emu_visible(self._h, p0.shape[0], _f(p0), _f(p1), out.ctypes.data_as(C.POINTER(C.c_uint8)))
        return out

    def primary_hits(self):
        n = self.width * self.height
        prim = np.zeros(n, np.uint32)
        tuv = np.zeros((n, 3), np.float32)
        lib().emu_primary_hits(self._h, prim.ctypes.data_as(C.POINTER(C.c_uint32)), _f(tuv))
        return prim.reshape(self.height, self.width), tuv.reshape(self.height, self.width, 3)

    def render(self, integ, spp, seed=0, rank=0, nranks=1):
        img = np.zeros((self.height, self.width, 3), np.float32)
        st = emu_stats()
        rc = lib().emu_render(self._h, C.byref(integ), spp, seed, rank, nranks, _f(img), C.byref(st))
        if rc != 0:
            raise ValueError(f"emu_render failed: {rc}")
        return img, st
