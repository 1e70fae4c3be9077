"""ctypes binding of the CPU oracle (oracle/_build/liboracle.so).

TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline and
--impl reference legs) may import this module; nothing under rustlight_b200/ does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from rustlight_b200 import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")

MATH_LIBM, MATH_SPEC = 0, 1
ACCEL_BVH, ACCEL_NAIVE = 0, 1
EST_GRAPH, EST_STREAM = 0, 1
SEED_PCG32, SEED_SPLITMIX64 = 0, 1


class orc_config(C.Structure):
    _fields_ = [("math_mode", C.c_uint32), ("accel_mode", C.c_uint32), ("estimator", C.c_uint32),
                ("seeding", C.c_uint32), ("nthreads", C.c_uint32), ("rank", C.c_uint32), ("nranks", C.c_uint32),
                ("sample_offset", C.c_uint32)]


class orc_stats(C.Structure):
    _fields_ = [("samples", C.c_uint64), ("segments", C.c_uint64), ("shadow_rays", C.c_uint64),
                ("shadow_visible", C.c_uint64), ("hits", C.c_uint64), ("max_depth_seen", C.c_uint64),
                ("nee_added", C.c_uint64), ("seconds", C.c_double), ("threads_used", C.c_uint32)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


def build(force=False):
    """Compile the oracle with oracle/Makefile (g++ -O2 -ffp-contract=off)."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "oracle.cpp")):
        subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
    return _SO


_lib = None
FP = C.POINTER(C.c_float)


def _f(a):
    return a.ctypes.data_as(FP)


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        build()
    L = C.CDLL(_SO)
    L.orc_scene_create.restype = C.c_void_p
    L.orc_scene_create.argtypes = [C.POINTER(_abi.rl_scene_desc), C.c_char_p, C.c_size_t]
    L.orc_scene_destroy.argtypes = [C.c_void_p]
    L.orc_render.argtypes = [C.c_void_p, C.POINTER(_abi.rl_integrator_desc), C.c_uint32, C.c_uint64, C.c_uint32,
                             C.POINTER(orc_config), FP, C.POINTER(orc_stats)]
    L.orc_trace.argtypes = [C.c_void_p, C.c_uint32, C.c_size_t, FP, FP, C.POINTER(C.c_uint32), FP, FP, FP, FP, FP]
    L.orc_visible.argtypes = [C.c_void_p, C.c_uint32, C.c_size_t, FP, FP, C.POINTER(C.c_uint8)]
    L.orc_primary_hits.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32), FP]
    L.orc_bvh_info.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), FP, FP]
    L.orc_intersect_tri.argtypes = [FP, FP, FP, FP, FP, FP, FP, FP, FP, FP]
    L.orc_aabb_intersect.argtypes = [FP, FP, FP, FP, C.c_float, C.c_float, FP]
    L.orc_frame.argtypes = [FP, FP]
    L.orc_cosine_sample_hemisphere.argtypes = [C.c_uint32, C.c_float, C.c_float, FP]
    L.orc_uniform_sample_triangle.argtypes = [C.c_float, C.c_float, FP]
    L.orc_dist1d_normalize.restype = C.c_float
    L.orc_dist1d_normalize.argtypes = [FP, C.c_uint32, FP]
    L.orc_dist1d_sample_discrete.restype = C.c_uint32
    L.orc_dist1d_sample_discrete.argtypes = [FP, C.c_uint32, C.c_float]
    L.orc_mis_weight.restype = C.c_float
    L.orc_mis_weight.argtypes = [C.c_float, C.c_float]
    L.orc_bsdf_sample.argtypes = [C.c_uint32, C.POINTER(_abi.rl_material), FP, C.c_float, C.c_float, FP, FP, FP]
    L.orc_bsdf_pdf.restype = C.c_float
    L.orc_bsdf_pdf.argtypes = [C.c_uint32, C.POINTER(_abi.rl_material), FP, FP]
    L.orc_bsdf_eval.argtypes = [C.c_uint32, C.POINTER(_abi.rl_material), FP, FP, FP]
    L.orc_sample_light.argtypes = [C.c_void_p, FP, C.c_float, C.c_float, C.c_float, C.c_float, FP, FP, FP, FP, FP]
    L.orc_direct_pdf.restype = C.c_float
    L.orc_direct_pdf.argtypes = [C.c_void_p, C.c_uint32, FP, FP, FP, FP]
    L.orc_camera_new.argtypes = [C.c_uint32, C.c_uint32, C.c_int, C.c_float, FP, C.c_int, FP]
    L.orc_camera_generate.argtypes = [C.c_void_p, C.c_float, C.c_float, FP, FP]
    L.orc_sampler_block_stream.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, FP]
    L.orc_sampler_counter.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, FP]
    L.orc_xoshiro_next_u64.restype = C.c_uint64
    L.orc_xoshiro_next_u64.argtypes = [C.POINTER(C.c_uint64)]
    L.orc_spec_sincos.argtypes = [C.c_float, FP, FP]
    L.orc_ats_sample.argtypes = [C.c_void_p, C.c_float, FP, FP, C.c_int, C.POINTER(C.c_uint32), FP]
    L.orc_ats_pdf.argtypes = [C.c_void_p, C.c_uint32, FP, FP, C.c_int, FP]
    L.orc_spec_atan2.restype = C.c_float
    L.orc_spec_atan2.argtypes = [C.c_float, C.c_float]
    L.orc_spec_acos.restype = C.c_float
    L.orc_spec_acos.argtypes = [C.c_float]
    L.orc_env_eval_pdf.argtypes = [C.c_void_p, C.c_uint32, FP, FP, FP]
    L.orc_env_sample.argtypes = [C.c_void_p, C.c_uint32, C.c_float, C.c_float, FP, FP, FP]
    L.orc_spec_powf.restype = C.c_float
    L.orc_spec_powf.argtypes = [C.c_float, C.c_float]
    L.orc_path_sample.argtypes = [C.c_void_p, C.POINTER(_abi.rl_integrator_desc), C.c_uint64, C.c_uint32, C.c_uint32,
                                  C.c_uint32, C.POINTER(orc_config), FP, C.POINTER(C.c_uint32),
                                  C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    _lib = L
    return L


def f32(x):
    return np.ascontiguousarray(x, dtype=np.float32)


def config(math_mode=MATH_SPEC, accel_mode=ACCEL_BVH, estimator=EST_GRAPH, seeding=SEED_PCG32, nthreads=0, rank=0,
           nranks=1, sample_offset=0):
    return orc_config(math_mode, accel_mode, estimator, seeding, nthreads, rank, nranks, sample_offset)


class OracleScene:
    """The oracle's Scene + BVHAccel + EmitterSampler for one rl_scene_desc."""

    def __init__(self, scene):
        self._scene = scene  # keep the host scene (and the buffers the desc points to) alive
        err = C.create_string_buffer(512)
        self._h = lib().orc_scene_create(scene.desc, err, 512)
        if not self._h:
            raise RuntimeError("oracle: " + err.value.decode())
        self.width, self.height = scene.size

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_scene_destroy(self._h)
            self._h = None

    def render(self, integ, spp, seed=0, sampler_mode=_abi.RL_SAMPLER_COUNTER, cfg=None):
        cfg = cfg or config()
        img = np.zeros((self.height, self.width, 3), dtype=np.float32)
        st = orc_stats()
        rc = lib().orc_render(self._h, C.byref(integ), spp, seed, sampler_mode, C.byref(cfg), _f(img), C.byref(st))
        if rc != 0:
            raise ValueError(f"orc_render failed: {rc}")
        return img, st

    def ats_sample(self, r, p, n=None):
        """LightSamplerATS::sample(r, importance_point(p, n)) -> (global triangle index, pdf)."""
        prim, pdf = C.c_uint32(), C.c_float()
        nn = np.zeros(3, np.float32) if n is None else np.ascontiguousarray(n, np.float32)
        if lib().orc_ats_sample(self._h, float(r), _f(np.ascontiguousarray(p, np.float32)), _f(nn), 0 if n is None else 1, C.byref(prim), C.byref(pdf)) != 0:
            raise ValueError("no light tree")
        return prim.value, pdf.value

    def ats_pdf(self, prim, p, n=None):
        pdf = C.c_float()
        nn = np.zeros(3, np.float32) if n is None else np.ascontiguousarray(n, np.float32)
        if lib().orc_ats_pdf(self._h, int(prim), _f(np.ascontiguousarray(p, np.float32)), _f(nn), 0 if n is None else 1, C.byref(pdf)) != 0:
            raise ValueError("no light tree / not a light")
        return pdf.value

    def env_eval_pdf(self, d, math_mode=MATH_SPEC):
        """EnvironmentLightColor::{eval, pdf} of the scene's environment for the direction d."""
        rgb, pdf = np.zeros(3, np.float32), C.c_float()
        if lib().orc_env_eval_pdf(self._h, math_mode, _f(np.ascontiguousarray(d, np.float32)), _f(rgb), C.byref(pdf)) != 0:
            raise ValueError("no environment")
        return rgb, pdf.value

    def env_sample(self, u0, u1, math_mode=MATH_SPEC):
        """EnvironmentLightColor::sample_direction((u0, u1)) -> (d, colour, pdf)."""
        d, rgb, pdf = np.zeros(3, np.float32), np.zeros(3, np.float32), C.c_float()
        if lib().orc_env_sample(self._h, math_mode, float(u0), float(u1), _f(d), _f(rgb), C.byref(pdf)) != 0:
            raise ValueError("no environment")
        return d, rgb, pdf.value

    def trace(self, o, d, accel_mode=ACCEL_BVH, full=False):
        o, d = f32(o).reshape(-1, 3), f32(d).reshape(-1, 3)
        n = o.shape[0]
        prim = np.zeros(n, dtype=np.uint32)
        tuv = np.zeros((n, 3), dtype=np.float32)
        extra = [np.zeros((n, 3), dtype=np.float32) for _ in range(4)] if full else [None] * 4
        lib().orc_trace(self._h, accel_mode, n, _f(o), _f(d), prim.ctypes.data_as(C.POINTER(C.c_uint32)), _f(tuv),
                        *[(_f(e) if e is not None else None) for e in extra])
        if full:
            return prim, tuv, dict(p=extra[0], n_g=extra[1], n_s=extra[2], wi=extra[3])
        return prim, tuv

    def visible(self, p0, p1, accel_mode=ACCEL_BVH):
        p0, p1 = f32(p0).reshape(-1, 3), f32(p1).reshape(-1, 3)
        out = np.zeros(p0.shape[0], dtype=np.uint8)
        lib().orc_visible(self._h, accel_mode, p0.shape[0], _f(p0), _f(p1), out.ctypes.data_as(C.POINTER(C.c_uint8)))
        return out

    def primary_hits(self, accel_mode=ACCEL_BVH):
        n = self.width * self.height
        prim = np.zeros(n, dtype=np.uint32)
        tuv = np.zeros((n, 3), dtype=np.float32)
        lib().orc_primary_hits(self._h, accel_mode, prim.ctypes.data_as(C.POINTER(C.c_uint32)), _f(tuv))
        return prim.reshape(self.height, self.width), tuv.reshape(self.height, self.width, 3)

    def bvh_info(self):
        nn, npr = C.c_uint32(), C.c_uint32()
        mn, mx = np.zeros(3, np.float32), np.zeros(3, np.float32)
        lib().orc_bvh_info(self._h, C.byref(nn), C.byref(npr), _f(mn), _f(mx))
        return dict(nnodes=nn.value, nprims=npr.value, root_min=mn, root_max=mx)

    def camera_generate(self, px, py):
        o, d = np.zeros(3, np.float32), np.zeros(3, np.float32)
        lib().orc_camera_generate(self._h, px, py, _f(o), _f(d))
        return o, d

    def sample_light(self, x, r_sel, r, u0, u1):
        x = f32(x)
        p, n, d, w = (np.zeros(3, np.float32) for _ in range(4))
        pdf = C.c_float()
        mesh = lib().orc_sample_light(self._h, _f(x), r_sel, r, u0, u1, _f(p), _f(n), _f(d), _f(w), C.byref(pdf))
        return dict(mesh=mesh, p=p, n=n, d=d, weight=w, pdf=pdf.value)

    def direct_pdf(self, mesh, o, p, n, dirv):
        o, p, n, dirv = f32(o), f32(p), f32(n), f32(dirv)
        return lib().orc_direct_pdf(self._h, mesh, _f(o), _f(p), _f(n), _f(dirv))

    def path_sample(self, integ, seed, px, py, sample, cfg=None):
        cfg = cfg or config()
        rgb = np.zeros(3, np.float32)
        a, b, c = C.c_uint32(), C.c_uint32(), C.c_uint32()
        lib().orc_path_sample(self._h, C.byref(integ), seed, px, py, sample, C.byref(cfg), _f(rgb), C.byref(a), C.byref(b),
                              C.byref(c))
        return rgb, a.value, b.value, c.value


# ---- unit-level wrappers ----------------------------------------------------------------------
def intersect_tri(v0, v1, v2, o, d, t_max=np.finfo(np.float32).max):
    v0, v1, v2, o, d = map(f32, (v0, v1, v2, o, d))
    t = C.c_float(t_max)
    u, v = C.c_float(), C.c_float()
    p, n = np.zeros(3, np.float32), np.zeros(3, np.float32)
    hit = lib().orc_intersect_tri(_f(v0), _f(v1), _f(v2), _f(o), _f(d), C.byref(t), C.byref(u), C.byref(v), _f(p), _f(n))
    return (bool(hit), t.value, u.value, v.value, p, n)


def aabb_intersect(pmin, pmax, o, d, tnear=1e-4, tfar=np.finfo(np.float32).max):
    pmin, pmax, o, d = map(f32, (pmin, pmax, o, d))
    t = C.c_float()
    hit = lib().orc_aabb_intersect(_f(pmin), _f(pmax), _f(o), _f(d), tnear, tfar, C.byref(t))
    return (bool(hit), t.value)


def frame(n):
    n = f32(n)
    out = np.zeros(9, np.float32)
    lib().orc_frame(_f(n), _f(out))
    return out.reshape(3, 3)  # rows: x, y, z axes


def cosine_sample_hemisphere(u0, u1, math_mode=MATH_SPEC):
    out = np.zeros(3, np.float32)
    lib().orc_cosine_sample_hemisphere(math_mode, u0, u1, _f(out))
    return out


def uniform_sample_triangle(u0, u1):
    out = np.zeros(2, np.float32)
    lib().orc_uniform_sample_triangle(u0, u1, _f(out))
    return out


def dist1d_normalize(elements):
    e = f32(elements)
    cdf = np.zeros(e.size + 1, np.float32)
    func_int = lib().orc_dist1d_normalize(_f(e), e.size, _f(cdf))
    return cdf, func_int


def dist1d_sample_discrete(cdf, v):
    cdf = f32(cdf)
    return int(lib().orc_dist1d_sample_discrete(_f(cdf), cdf.size, v))


def mis_weight(a, b):
    return lib().orc_mis_weight(a, b)


def bsdf_sample(mat, wi, s0, s1, math_mode=MATH_SPEC):
    wi = f32(wi)
    w, d = np.zeros(3, np.float32), np.zeros(3, np.float32)
    pdf = C.c_float()
    ok = lib().orc_bsdf_sample(math_mode, C.byref(mat), _f(wi), s0, s1, _f(w), _f(d), C.byref(pdf))
    return (bool(ok), w, d, pdf.value)


def bsdf_sample_ex(mat, wi, s0, s1, math_mode=MATH_SPEC):
    """(ok, weight, d, pdf, discrete): discrete <=> the sampled pdf is PDF::Discrete."""
    wi = f32(wi)
    w, d = np.zeros(3, np.float32), np.zeros(3, np.float32)
    pdf = C.c_float()
    rc = lib().orc_bsdf_sample(math_mode, C.byref(mat), _f(wi), s0, s1, _f(w), _f(d), C.byref(pdf))
    return (rc != 0, w, d, pdf.value, rc == 2)


def bsdf_flags(mat):
    f = lib().orc_bsdf_flags(C.byref(mat))
    return dict(twosided=bool(f & 1), smooth=bool(f & 2))


def bsdf_pdf(mat, wi, wo, math_mode=MATH_SPEC):
    wi, wo = f32(wi), f32(wo)
    return lib().orc_bsdf_pdf(math_mode, C.byref(mat), _f(wi), _f(wo))


def bsdf_eval(mat, wi, wo, math_mode=MATH_SPEC):
    wi, wo = f32(wi), f32(wo)
    out = np.zeros(3, np.float32)
    lib().orc_bsdf_eval(math_mode, C.byref(mat), _f(wi), _f(wo), _f(out))
    return out


def camera_new(w, h, fov_deg, to_world, fov_axis="y", flip=False):
    tw = f32(to_world).ravel()
    out = np.zeros(16, np.float32)
    if lib().orc_camera_new(w, h, 1 if fov_axis == "x" else 0, fov_deg, _f(tw), 1 if flip else 0, _f(out)) != 0:
        raise RuntimeError("orc_camera_new failed")
    return out


def sampler_block_stream(seed, block, n, seeding=SEED_PCG32):
    out = np.zeros(n, np.float32)
    lib().orc_sampler_block_stream(seed, seeding, block, n, _f(out))
    return out


def sampler_counter(seed, pixel, sample, n):
    out = np.zeros(n, np.float32)
    lib().orc_sampler_counter(seed, pixel, sample, n, _f(out))
    return out


def xoshiro_next_u64(state):
    st = (C.c_uint64 * 4)(*state)
    r = lib().orc_xoshiro_next_u64(st)
    return r, list(st)


def spec_atan2(y, x):
    return lib().orc_spec_atan2(float(y), float(x))


def spec_acos(x):
    return lib().orc_spec_acos(float(x))


def spec_sincos(x):
    s, c = C.c_float(), C.c_float()
    lib().orc_spec_sincos(x, C.byref(s), C.byref(c))
    return s.value, c.value


def spec_powf(x, y):
    return lib().orc_spec_powf(x, y)
