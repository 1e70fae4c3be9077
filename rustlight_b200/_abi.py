"""ctypes mirror of include/rl_b200.h (struct layouts and enum values only)."""
import ctypes as C

RL_OK = 0
RL_ERR_INVALID = -1
RL_ERR_CUDA = -2
RL_ERR_NCCL = -3
RL_ERR_UNSUPPORTED = -4
RL_ERR_NOMEM = -5

RL_BSDF_DIFFUSE = 0
RL_BSDF_PHONG = 1
RL_BSDF_METAL = 2
RL_BSDF_GLASS = 3
RL_BSDF_SUBSTRATE = 4
RL_BSDF_BLEND = 5
RL_MICROFACET_NONE = 0
RL_MICROFACET_GGX = 1
RL_MICROFACET_BECKMANN = 2

RL_INTEGRATOR_PATH = 0
RL_INTEGRATOR_DIRECT = 1
RL_INTEGRATOR_AO = 2

RL_STRATEGY_ALL = 0
RL_STRATEGY_BSDF = 1
RL_STRATEGY_EMITTER = 2

RL_SAMPLER_BLOCK_STREAM = 0
RL_SAMPLER_COUNTER = 1

RL_MISS = 0xFFFFFFFF


class rl_material(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("kd", C.c_float * 3), ("ks", C.c_float * 3),
                ("exponent", C.c_float), ("weight_specular", C.c_float), ("kt", C.c_float * 3),
                ("eta", C.c_float * 3), ("k", C.c_float * 3), ("ior", C.c_float), ("alpha", C.c_float),
                ("microfacet", C.c_uint32), ("kd_texture", C.c_uint32), ("ks_texture", C.c_uint32),
                ("kt_texture", C.c_uint32), ("eta_texture", C.c_uint32), ("k_texture", C.c_uint32),
                ("blend_a", C.c_uint32), ("blend_b", C.c_uint32), ("blend_weight", C.c_float)]


class rl_mesh_desc(C.Structure):
    _fields_ = [("P", C.POINTER(C.c_float)), ("nverts", C.c_uint32),
                ("idx", C.POINTER(C.c_uint32)), ("ntris", C.c_uint32),
                ("N", C.POINTER(C.c_float)), ("UV", C.POINTER(C.c_float)),
                ("mat", rl_material), ("emission_kind", C.c_uint32), ("emission", C.c_float * 3), ("emission_texture", C.c_uint32)]


RL_EMISSION_ZERO, RL_EMISSION_COLOR, RL_EMISSION_HSV, RL_EMISSION_TEXTURE = 0, 1, 2, 3  # rl_emission_kind (geometry.rs:99-104)
RL_TEX_BITMAP = 1
RL_TEX_CHECKERBOARD = 2
RL_TEX_GRID = 3


class rl_texture(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("width", C.c_uint32), ("height", C.c_uint32), ("pixels", C.POINTER(C.c_float)),
                ("color0", C.c_float * 3), ("color1", C.c_float * 3), ("line_width", C.c_float),
                ("offset", C.c_float * 2), ("scale", C.c_float * 2)]


RL_LIGHT_POINT = 0
RL_LIGHT_DIRECTIONAL = 1


class rl_light_desc(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("intensity", C.c_float * 3), ("v", C.c_float * 3)]


class rl_camera_desc(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32),
                ("sample_to_camera", C.c_float * 16), ("to_world", C.c_float * 16)]


class rl_scene_desc(C.Structure):
    _fields_ = [("nmeshes", C.c_uint32), ("meshes", C.POINTER(rl_mesh_desc)),
                ("camera", rl_camera_desc), ("has_volume", C.c_uint32), ("has_environment", C.c_uint32),
                ("nlights", C.c_uint32), ("lights", C.POINTER(rl_light_desc)),
                ("ntextures", C.c_uint32), ("textures", C.POINTER(rl_texture)), ("environment", C.c_float * 3),
                ("nsubmaterials", C.c_uint32), ("submaterials", C.POINTER(rl_material)), ("environment_texture", C.c_uint32), ("use_ats", C.c_uint32)]


class rl_integrator_desc(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("min_depth", C.c_int32), ("max_depth", C.c_int32),
                ("rr_depth", C.c_int32), ("strategy", C.c_uint32), ("single_scattering", C.c_uint32),
                ("nb_bsdf_samples", C.c_uint32), ("nb_light_samples", C.c_uint32),
                ("ao_max_distance", C.c_float), ("ao_normal_correction", C.c_uint32)]


class rl_render_opts(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("spp", C.c_uint32), ("seed", C.c_uint64),
                ("sampler_mode", C.c_uint32), ("batch_spp", C.c_uint32),
                ("material_sort", C.c_uint32), ("sample_offset", C.c_uint32)]


class rl_stats(C.Structure):
    _fields_ = [("samples", C.c_uint64), ("segments", C.c_uint64), ("shadow_rays", C.c_uint64),
                ("shadow_visible", C.c_uint64), ("hits", C.c_uint64), ("max_depth_seen", C.c_uint64),
                ("kernel_launches", C.c_uint64), ("ms_total", C.c_double), ("ms_raygen", C.c_double),
                ("ms_trace", C.c_double), ("ms_shade", C.c_double), ("ms_shadow", C.c_double),
                ("ms_accum", C.c_double), ("ms_h2d", C.c_double), ("ms_d2h", C.c_double),
                ("ms_reduce", C.c_double), ("ms_tail", C.c_double), ("shadow_traced", C.c_uint64),
                ("launches_trace", C.c_uint64), ("launches_shade", C.c_uint64)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class rl_bvh_info(C.Structure):
    _fields_ = [("ntris", C.c_uint32), ("nnodes", C.c_uint32), ("nleaves", C.c_uint32),
                ("max_depth", C.c_uint32), ("root_min", C.c_float * 3), ("root_max", C.c_float * 3),
                ("smem_resident", C.c_uint32), ("flat_groups", C.c_uint32), ("flat_pairs", C.c_uint32),
                ("flat_singles", C.c_uint32), ("flat_delta", C.c_float)]


class rl_layout_info(C.Structure):
    _fields_ = [("ray_bytes", C.c_uint32), ("hit_bytes", C.c_uint32), ("state_bytes", C.c_uint32),
                ("shadow_bytes", C.c_uint32), ("accum_bytes", C.c_uint32),
                ("max_paths_in_flight", C.c_uint64)]


def path_desc(min_depth=0, max_depth=None, rr_depth=0, strategy=RL_STRATEGY_ALL, single_scattering=False):
    """IntegratorPathTracing with the CLI defaults of examples/cli.rs:54-61,167."""
    opt = lambda v: -1 if v is None else int(v)
    return rl_integrator_desc(RL_INTEGRATOR_PATH, opt(min_depth), opt(max_depth), opt(rr_depth),
                              strategy, 1 if single_scattering else 0, 1, 1)


def ao_desc(max_distance=1.0, normal_correction=False):
    """IntegratorAO with the CLI defaults of examples/cli.rs:150-155 (`-d inf` = None)."""
    return rl_integrator_desc(RL_INTEGRATOR_AO, 0, -1, 0, RL_STRATEGY_ALL, 0, 1, 0,
                              -1.0 if max_distance is None else float(max_distance), 1 if normal_correction else 0)


def direct_desc(nb_bsdf_samples=1, nb_light_samples=1):
    """IntegratorDirect with the CLI defaults of examples/cli.rs:157-160."""
    return rl_integrator_desc(RL_INTEGRATOR_DIRECT, 0, -1, 0, RL_STRATEGY_ALL, 0,
                              int(nb_bsdf_samples), int(nb_light_samples))
