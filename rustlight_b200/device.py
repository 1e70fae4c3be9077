"""Python face of the CUDA library (librl_b200.so, C ABI in include/rl_b200.h).

Names mirror the reference: `IntegratorPathTracing` / `IntegratorDirect` with
`compute(sampler, scene) -> BufferCollection` (src/integrators/mod.rs:219-228,
explicit/path.rs:14-20, direct.rs:5-8), `IndependentSampler(seed)` for `-r independent:<seed>`
(examples/cli.rs:886-890), `Acceleration.trace/visible` (src/accel.rs:9-12).

There is no CPU fallback: loading fails loudly when the library is missing and `Context()`
fails loudly when no GPU is present.
"""
import ctypes as C
import os

import numpy as np

from . import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None
FP = C.POINTER(C.c_float)
U32P = C.POINTER(C.c_uint32)


class DeviceError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"rl_b200 error {code}: {msg}")
        self.code = code


def lib():
    global _lib
    if _lib is None:
        path = os.environ.get("RL_B200_LIB") or os.path.join(_HERE, "librl_b200.so")  # override: kernel A/B experiments only
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: the CUDA extension is not built "
                               "(python -c 'import __graft_entry__ as g; g.build()'); there is no CPU fallback")
        L = C.CDLL(path)
        L.rl_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.rl_destroy.argtypes = [C.c_void_p]
        L.rl_nccl_unique_id.argtypes = [C.c_void_p]
        L.rl_last_error.restype = C.c_char_p
        L.rl_last_error.argtypes = [C.c_void_p]
        L.rl_set_profiling.argtypes = [C.c_void_p, C.c_int]
        L.rl_scene_create.argtypes = [C.c_void_p, C.POINTER(_abi.rl_scene_desc), C.POINTER(C.c_void_p)]
        L.rl_scene_destroy.argtypes = [C.c_void_p, C.c_void_p]
        L.rl_scene_bvh_info.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(_abi.rl_bvh_info)]
        L.rl_render.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(_abi.rl_integrator_desc),
                                C.POINTER(_abi.rl_render_opts), FP, C.POINTER(_abi.rl_stats)]
        L.rl_render_device.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(_abi.rl_integrator_desc),
                                       C.POINTER(_abi.rl_render_opts), C.c_void_p, C.POINTER(_abi.rl_stats)]
        L.rl_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, FP, FP, U32P, FP]
        L.rl_visible.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, FP, FP, C.POINTER(C.c_uint8)]
        L.rl_primary_hits.argtypes = [C.c_void_p, C.c_void_p, U32P, FP]
        L.rl_layout.argtypes = [C.c_void_p, C.POINTER(_abi.rl_layout_info)]
        L.rl_host_alloc.restype = C.c_void_p
        L.rl_host_alloc.argtypes = [C.c_size_t]
        L.rl_host_free.argtypes = [C.c_void_p]
        _lib = L
    return _lib


class PinnedImage:
    """H x W x 3 float32 frame in page-locked host memory (rl_host_alloc): `.array` is a numpy view, valid until close()."""

    def __init__(self, h, w):
        import numpy as np
        nbytes = h * w * 3 * 4
        self._p = lib().rl_host_alloc(nbytes)
        if not self._p:
            raise RuntimeError("rl_host_alloc failed")
        self.array = np.ctypeslib.as_array((C.c_float * (h * w * 3)).from_address(self._p)).reshape(h, w, 3)
        self.array[...] = 0.0

    def close(self):
        if self._p:
            self.array = None
            lib().rl_host_free(self._p)
            self._p = None


def nccl_unique_id():
    buf = C.create_string_buffer(128)
    rc = lib().rl_nccl_unique_id(buf)
    if rc != 0:
        raise DeviceError(rc, lib().rl_last_error(None).decode())
    return buf.raw


class Context:
    """One GPU (one rank).  Replaces the rayon pool of integrators/mod.rs:452-459."""

    def __init__(self, device=0, nranks=1, rank=0, nccl_id=None):
        self._h = C.c_void_p()
        rc = lib().rl_create(C.byref(self._h), device, nranks, rank, nccl_id)
        if rc != 0:
            self._h = None
            raise DeviceError(rc, lib().rl_last_error(None).decode())
        self.device, self.nranks, self.rank = device, nranks, rank

    def close(self):
        if getattr(self, "_h", None):
            lib().rl_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def _check(self, rc):
        if rc != 0:
            raise DeviceError(rc, lib().rl_last_error(self._h).decode())

    def set_profiling(self, on):
        self._check(lib().rl_set_profiling(self._h, int(on)))  # False/0 off, True/1 per-stage kernels, 2 the kernels of an untimed frame

    def layout(self):
        li = _abi.rl_layout_info()
        self._check(lib().rl_layout(self._h, C.byref(li)))
        return li


class DeviceScene:
    """Device-resident scene + LBVH (`Acceleration`).  Replaces BVHAccel::new + build_emitters."""

    def __init__(self, ctx, scene):
        self.ctx = ctx
        self.host_scene = scene
        self.width, self.height = scene.size
        self._h = C.c_void_p()
        ctx._check(lib().rl_scene_create(ctx._h, scene.desc, C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) and getattr(self.ctx, "_h", None):
            lib().rl_scene_destroy(self.ctx._h, self._h)
        self._h = None

    def __del__(self):
        self.close()

    def bvh_info(self):
        bi = _abi.rl_bvh_info()
        self.ctx._check(lib().rl_scene_bvh_info(self.ctx._h, self._h, C.byref(bi)))
        return bi

    # Acceleration::trace
    def trace(self, o, d):
        o = np.ascontiguousarray(o, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(d, np.float32).reshape(-1, 3)
        prim = np.zeros(o.shape[0], np.uint32)
        tuv = np.zeros((o.shape[0], 3), np.float32)
        self.ctx._check(lib().rl_trace(self.ctx._h, self._h, o.shape[0], o.ctypes.data_as(FP), d.ctypes.data_as(FP),
                                       prim.ctypes.data_as(U32P), tuv.ctypes.data_as(FP)))
        return prim, tuv

    # Acceleration::visible
    def visible(self, p0, p1):
        p0 = np.ascontiguousarray(p0, np.float32).reshape(-1, 3)
        p1 = np.ascontiguousarray(p1, np.float32).reshape(-1, 3)
        out = np.zeros(p0.shape[0], np.uint8)
        self.ctx._check(lib().rl_visible(self.ctx._h, self._h, p0.shape[0], p0.ctypes.data_as(FP),
                                         p1.ctypes.data_as(FP), out.ctypes.data_as(C.POINTER(C.c_uint8))))
        return out

    def primary_hits(self):
        n = self.width * self.height
        prim = np.zeros(n, np.uint32)
        tuv = np.zeros((n, 3), np.float32)
        self.ctx._check(lib().rl_primary_hits(self.ctx._h, self._h, prim.ctypes.data_as(U32P), tuv.ctypes.data_as(FP)))
        return prim.reshape(self.height, self.width), tuv.reshape(self.height, self.width, 3)

    def render(self, integ, spp, seed=0, out=None, batch_spp=0, material_sort=2, device_out=None, want_image=True, sample_offset=0):
        """rl_render: returns (image HxWx3 float32 or None, rl_stats)."""
        opts = _abi.rl_render_opts(C.sizeof(_abi.rl_render_opts), int(spp), int(seed), _abi.RL_SAMPLER_COUNTER,
                                   int(batch_spp), int(material_sort), int(sample_offset))
        st = _abi.rl_stats()
        if device_out is not None:
            self.ctx._check(lib().rl_render_device(self.ctx._h, self._h, C.byref(integ), C.byref(opts),
                                                   C.c_void_p(device_out), C.byref(st)))
            return None, st
        if out is None and want_image:
            out = np.zeros((self.height, self.width, 3), np.float32)
        ptr = out.ctypes.data_as(FP) if out is not None else None
        self.ctx._check(lib().rl_render(self.ctx._h, self._h, C.byref(integ), C.byref(opts), ptr, C.byref(st)))
        return out, st


class IndependentSampler:
    """`-r independent:<seed>`; on the GPU the stream is the counter-based mode B."""

    def __init__(self, seed=0):
        self.seed = int(seed)
        self.passes = 0  # compute() calls so far: pass p renders samples [p * spp, (p + 1) * spp), like rl_integrators.hpp


class BufferCollection:
    """integrators/mod.rs:48-52: named bitmaps; this path only produces "primal"."""

    def __init__(self, primal, stats=None):
        self.values = {"primal": primal}
        self.stats = stats

    def save(self, name, filename):
        """BufferCollection::save -> Bitmap::save (structure.rs:528-545): .pfm or .png by extension."""
        from .host import save_image
        save_image(filename, self.values[name])


class _IntegratorBase:
    def compute(self, sampler, scene, ctx=None, **kw):
        """Integrator::compute: `scene` is a DeviceScene (or a host Scene + ctx); spp = scene.nb_samples."""
        if not isinstance(scene, DeviceScene):
            scene = DeviceScene(ctx or Context(), scene)
        spp = kw.pop("spp", None) or scene.host_scene.nb_samples
        # the reference's master sampler moves on with every compute() (generate_img_blocks clones it per block, mod.rs:351-374), so the
        # passes of an averaging wrapper (avg.rs:45-65) never repeat a sample; here the sample offset advances instead
        kw.setdefault("sample_offset", getattr(sampler, "passes", 0) * spp)
        img, st = scene.render(self.desc(), spp, seed=sampler.seed, **kw)
        if hasattr(sampler, "passes"):
            sampler.passes += 1
        return BufferCollection(img, st)


class IntegratorPathTracing(_IntegratorBase):
    """explicit/path.rs:14-20 with the CLI defaults (cli.rs:54-61,167)."""

    def __init__(self, min_depth=0, max_depth=None, rr_depth=0, strategy="all", single_scattering=False):
        self.min_depth, self.max_depth, self.rr_depth = min_depth, max_depth, rr_depth
        self.strategy = {"all": _abi.RL_STRATEGY_ALL, "bsdf": _abi.RL_STRATEGY_BSDF,
                         "emitter": _abi.RL_STRATEGY_EMITTER}[strategy] if isinstance(strategy, str) else strategy
        self.single_scattering = single_scattering

    def desc(self):
        return _abi.path_desc(self.min_depth, self.max_depth, self.rr_depth, self.strategy, self.single_scattering)


class IntegratorAO(_IntegratorBase):
    """ao.rs:4-7 with the CLI defaults (cli.rs:150-155); max_distance=None is `-d inf`."""

    def __init__(self, max_distance=1.0, normal_correction=False):
        self.max_distance, self.normal_correction = max_distance, normal_correction

    def desc(self):
        return _abi.ao_desc(self.max_distance, self.normal_correction)


class IntegratorDirect(_IntegratorBase):
    """direct.rs:5-8 with the CLI defaults (cli.rs:157-160)."""

    def __init__(self, nb_bsdf_samples=1, nb_light_samples=1):
        self.nb_bsdf_samples, self.nb_light_samples = nb_bsdf_samples, nb_light_samples

    def desc(self):
        return _abi.direct_desc(self.nb_bsdf_samples, self.nb_light_samples)
