/*
 * rl_b200.h -- C ABI of the B200 wavefront path tracer that sits behind rustlight's
 * `path` and `direct` integrators.
 *
 * The reference (beltegeuse/rustlight @ 864df34) has no FFI; its seam for this path is the
 * pair of Rust traits
 *     Integrator::compute(&mut self, &mut dyn Sampler, &dyn Acceleration, &Scene)
 *         -> BufferCollection                         src/integrators/mod.rs:219-228
 *     Acceleration::{trace, visible}                  src/accel.rs:9-12
 * as implemented by `compute_mc` (src/integrators/mod.rs:403-450) for
 * IntegratorPathTracing (src/integrators/explicit/path.rs:187-238) and
 * IntegratorDirect (src/integrators/direct.rs:10-233).  Every entry point below names the
 * reference item it replaces.  Plain pointers and sizes only; all host buffers are owned by
 * the caller, all device memory by the library.  Nothing here aborts the process: every call
 * returns RL_OK (0) or a negative rl_status, and rl_last_error() gives the text (the
 * reference panics instead: e.g. src/integrators/mod.rs:410, src/structure.rs:707).
 *
 * Matrices are column-major f32[16] like cgmath::Matrix4 (m[4*col + row]).
 * Images are row-major, index = y*width + x, 3 floats (r,g,b) per pixel, like
 * Bitmap.colors (src/structure.rs:383-402).
 */
#ifndef RL_B200_H
#define RL_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RL_B200_ABI_VERSION 6

typedef struct rl_ctx rl_ctx;     /* one per GPU / per rank; single-owner, not thread-safe */
typedef struct rl_scene rl_scene; /* device-resident scene: geometry, LBVH, emitters, camera */

typedef enum rl_status {
    RL_OK = 0,
    RL_ERR_INVALID = -1,     /* bad argument (zero spp, null pointer, bad index, ...)          */
    RL_ERR_CUDA = -2,        /* a CUDA runtime call failed                                      */
    RL_ERR_NCCL = -3,        /* NCCL missing or a collective failed                             */
    RL_ERR_UNSUPPORTED = -4, /* feature outside the hot path (volume, env map, sampler mode A)  */
    RL_ERR_NOMEM = -5
} rl_status;

/* ---- materials: src/bsdfs/mod.rs:163-199 (trait BSDF) ------------------------------------ */
typedef enum rl_bsdf_kind {
    RL_BSDF_DIFFUSE = 0,   /* BSDFDiffuse   src/bsdfs/diffuse.rs:6-87                                     */
    RL_BSDF_PHONG = 1,     /* BSDFPhong     src/bsdfs/phong.rs:6-136                                      */
    RL_BSDF_METAL = 2,     /* BSDFMetal     src/bsdfs/metal.rs:6-177 (microfacet == NONE: pbrt "mirror")  */
    RL_BSDF_GLASS = 3,     /* BSDFGlass     src/bsdfs/glass.rs:35-192 (DELTA, not two-sided)              */
    RL_BSDF_SUBSTRATE = 4, /* BSDFSubstrate src/bsdfs/substrate.rs:8-225                                  */
    RL_BSDF_BLEND = 5      /* BSDFBlend     src/bsdfs/blend.rs:3-95: weight * bsdf1 + (1 - weight) * bsdf2; both rough (the reference asserts
                              !is_smooth()), two-sided, constant colours, not themselves blends                                       */
} rl_bsdf_kind;
typedef enum rl_microfacet { /* MicrofacetDistributionBSDF, src/bsdfs/distribution.rs:5-17; isotropic only (:62) */
    RL_MICROFACET_NONE = 0, /* distribution: None -> pure specular lobe (BSDFType::DELTA)                */
    RL_MICROFACET_GGX = 1,  /* what the PBRT route builds (bsdfs/mod.rs:260-291)                          */
    RL_MICROFACET_BECKMANN = 2
} rl_microfacet;

/* BSDFColor (bsdfs/mod.rs:11-101) other than Constant, for any colour slot of a material: rl_material.<slot>_texture = 1 + index
 * into rl_scene_desc.textures (0 = BSDFColor::Constant(<slot>)).  A mesh without uv coordinates evaluates a texture to black
 * ("Found a texture but no uv coordinate given", bsdfs/mod.rs:36-39).  BSDFPhong's weight_specular stays the constant the
 * caller computed (the reference computes it once from constant colours too, bsdfs/mod.rs:518-523). */
typedef enum rl_texture_kind {
    RL_TEX_BITMAP = 1,       /* nearest texel, Bitmap::pixel_uv (structure.rs:434-453)                         */
    RL_TEX_CHECKERBOARD = 2, /* bsdfs/mod.rs:43-65                                                             */
    RL_TEX_GRID = 3          /* bsdfs/mod.rs:66-99 (incl. its `uv.y + scale.y`)                                */
} rl_texture_kind;
typedef struct rl_texture {
    uint32_t kind;
    uint32_t width, height;     /* bitmap                                                                   */
    const float *pixels;        /* bitmap: width*height*3 RGB floats, Bitmap.colors order (index y*width+x) */
    float color0[3], color1[3]; /* checkerboard / grid                                                      */
    float line_width;           /* grid                                                                     */
    float offset[2], scale[2];  /* checkerboard / grid                                                      */
} rl_texture;

/* Other colours are BSDFColor::Constant.  Fields used per kind:
 *   DIFFUSE   kd
 *   PHONG     kd, ks, exponent, weight_specular
 *   METAL     ks = specular, eta, k, microfacet, alpha
 *   GLASS     ks = specular_reflectance, kt = specular_transmittance, ior = int_ior / ext_ior (BSDFGlass::eta)
 *   SUBSTRATE kd = diffuse, ks = specular, microfacet, alpha */
typedef struct rl_material {
    uint32_t kind;         /* rl_bsdf_kind                                                    */
    float kd[3];           /* diffuse reflectance                                             */
    float ks[3];           /* specular reflectance                                            */
    float exponent;        /* Phong exponent                                                  */
    float weight_specular; /* lum(ks)/(lum(kd)+lum(ks)), src/bsdfs/mod.rs:518-523             */
    float kt[3];           /* glass: specular transmittance                                   */
    float eta[3], k[3];    /* metal: real and imaginary part of the index of refraction       */
    float ior;             /* glass: relative index of refraction (!= 0, glass.rs:43-48)      */
    float alpha;           /* microfacet roughness alpha_u == alpha_v                          */
    uint32_t microfacet;   /* rl_microfacet                                                   */
    uint32_t kd_texture;   /* 0 = constant kd, else 1 + index into rl_scene_desc.textures     */
    /* the other colour slots bsdf_pbrt runs through bsdf_texture_match_pbrt (bsdfs/mod.rs:294-386): same encoding            */
    uint32_t ks_texture;   /* phong / substrate Ks, metal specular (mirror Kr), glass Kr      */
    uint32_t kt_texture;   /* glass Kt                                                        */
    uint32_t eta_texture;  /* metal eta                                                       */
    uint32_t k_texture;    /* metal k                                                         */
    uint32_t blend_a, blend_b; /* BLEND: bsdf1, bsdf2 as 1 + index into rl_scene_desc.submaterials */
    float blend_weight;        /* BLEND: BSDFBlend.weight in [0, 1]                                */
} rl_material;

/* ---- geometry: Mesh, src/geometry.rs:107-119 ---------------------------------------------- */
/* EmissionType (geometry.rs:99-104) and Mesh::emit(uv) (:184-206).  HSV and TEXTURE are what the CLI's `-x hvs-light` /
 * `-x texture-light` turn every mesh light into (examples/cli.rs:410-429, scale = luminance of its colour); they need uv
 * coordinates on the mesh (the reference unwraps them).  Mesh::flux uses Color::value(scale) for both (emitter.rs:591-599). */
typedef enum rl_emission_kind {
    RL_EMISSION_ZERO = 0,
    RL_EMISSION_COLOR = 1,  /* emit = v                                                                       */
    RL_EMISSION_HSV = 2,    /* emit = (x, 1 - x, 0) * scale with x = |uv.x| % 1 (a red-to-green ramp; sic)    */
    RL_EMISSION_TEXTURE = 3 /* emit = img.pixel_uv(uv) * scale                                                */
} rl_emission_kind;
typedef struct rl_mesh_desc {
    const float *P;      /* 3*nverts                                                          */
    uint32_t nverts;
    const uint32_t *idx; /* 3*ntris, indices into P/N/UV                                      */
    uint32_t ntris;
    const float *N;  /* 3*nverts shading normals or NULL (Mesh.normals: Option)               */
    const float *UV; /* 2*nverts or NULL (Mesh.uv: Option)                                    */
    rl_material mat;
    uint32_t emission_kind; /* rl_emission_kind                                               */
    float emission[3];      /* COLOR: v; HSV / TEXTURE: emission[0] = scale                   */
    uint32_t emission_texture; /* TEXTURE: 1 + index into rl_scene_desc.textures (a RL_TEX_BITMAP) */
} rl_mesh_desc;

/* ---- camera: src/camera.rs:5-15; matrices as built by Camera::new (camera.rs:31-67) ------- */
typedef struct rl_camera_desc {
    uint32_t width, height;
    float sample_to_camera[16];
    float to_world[16];
} rl_camera_desc;

/* ---- non-mesh emitters: what PBRTSceneLoader turns `LightSource "point" / "distant"` into (scene_loader.rs:207-240) -- */
typedef enum rl_light_kind {
    RL_LIGHT_POINT = 0,      /* PointEmitter      src/emitter.rs:186-250: intensity, position       */
    RL_LIGHT_DIRECTIONAL = 1 /* DirectionalLight  src/emitter.rs:96-190:  intensity, direction (light -> world, unit);
                                its bounding sphere comes from Scene::build_emitters (scene.rs:54-60), radius x 1.1 */
} rl_light_kind;
typedef struct rl_light_desc {
    uint32_t kind;
    float intensity[3];
    float v[3]; /* position (point) or direction (directional) */
} rl_light_desc;

/* ---- scene: src/scene.rs:16-30 ------------------------------------------------------------ */
typedef struct rl_scene_desc {
    uint32_t nmeshes;
    const rl_mesh_desc *meshes;
    rl_camera_desc camera;
    uint32_t has_volume;      /* must be 0: scene.volume == None on this path                 */
    uint32_t has_environment; /* 0: emitter_environment == None; 1: EnvironmentLight with EnvironmentLightColor::Constant(
                                 environment); 2: EnvironmentLightColor::Texture { image = textures[environment_texture - 1] }
                                 (lat-long map, importance-sampled through its Distribution2D; emitter.rs:300-568)           */
    uint32_t nlights;         /* Scene.emitters (EmittersState::Unbuild): sampled after the mesh lights, in this order */
    const rl_light_desc *lights;
    uint32_t ntextures;
    const rl_texture *textures;
    float environment[3];     /* constant environment radiance when has_environment == 1 */
    uint32_t nsubmaterials;   /* the bsdf1 / bsdf2 of RL_BSDF_BLEND materials */
    const rl_material *submaterials;
    uint32_t environment_texture; /* has_environment == 2: 1 + index into textures[] of an RL_TEX_BITMAP (EnvironmentLightColor::new_texture) */
    uint32_t use_ats;             /* Scene::build_emitters(true) (`-x ats`, examples/cli.rs:325,432): light sampling through the light tree
                                     LightSamplerATS (emitter.rs:1130-1400).  Mesh emitters only (the reference asserts is_surface()); `direct` only
                                     with nb_bsdf_samples == 0 */
} rl_scene_desc;

/* ---- integrators --------------------------------------------------------------------------- */
typedef enum rl_integrator_kind {
    RL_INTEGRATOR_PATH = 0,  /* IntegratorPathTracing  src/integrators/explicit/path.rs:14-20 */
    RL_INTEGRATOR_DIRECT = 1, /* IntegratorDirect      src/integrators/direct.rs:5-8          */
    RL_INTEGRATOR_AO = 2      /* IntegratorAO          src/integrators/ao.rs:4-7              */
} rl_integrator_kind;

typedef enum rl_path_strategy { /* IntegratorPathTracingStrategies, path.rs:9-13 */
    RL_STRATEGY_ALL = 0,
    RL_STRATEGY_BSDF = 1,
    RL_STRATEGY_EMITTER = 2
} rl_path_strategy;

typedef struct rl_integrator_desc {
    uint32_t kind;              /* rl_integrator_kind                                          */
    int32_t min_depth;          /* Option<u32>: -1 = None   (CLI default 0,  cli.rs:57-58)     */
    int32_t max_depth;          /* Option<u32>: -1 = None   (CLI default inf, cli.rs:55-56)    */
    int32_t rr_depth;           /* Option<u32>: -1 = None   (CLI default 0,  cli.rs:59-60)     */
    uint32_t strategy;          /* rl_path_strategy                                            */
    uint32_t single_scattering; /* path.rs:19                                                  */
    uint32_t nb_bsdf_samples;   /* direct.rs:6, CLI default 1                                  */
    uint32_t nb_light_samples;  /* direct.rs:7, CLI default 1                                  */
    float ao_max_distance;      /* ao.rs:5 Option<f32>: < 0 = None (`-d inf`); CLI default 1.0 */
    uint32_t ao_normal_correction; /* ao.rs:6                                                  */
} rl_integrator_desc;

/* Sampler streams.  Mode A is the reference's own: one sequential SmallRng per 16x16 block,
 * cloned in x-major block order (integrators/mod.rs:351-374, samplers/independent.rs:18-22);
 * it serialises 256*spp samples per stream and exists only in the CPU oracle.  Mode B is a
 * counter-based stream per (pixel, sample): same call sequence (next/next2d), different
 * numbers, independent of how the image is partitioned.  The GPU implements mode B. */
typedef enum rl_sampler_mode { RL_SAMPLER_BLOCK_STREAM = 0, RL_SAMPLER_COUNTER = 1 } rl_sampler_mode;

typedef struct rl_render_opts {
    uint32_t struct_size;   /* sizeof(rl_render_opts), for forward compatibility               */
    uint32_t spp;           /* scene.nb_samples; 0 is an error (integrators/mod.rs:410)        */
    uint64_t seed;          /* `-r independent:<seed>` (cli.rs:886-890)                        */
    uint32_t sampler_mode;  /* rl_sampler_mode; only RL_SAMPLER_COUNTER on the GPU             */
    uint32_t batch_spp;     /* samples per pixel in flight per wavefront batch; 0 = auto       */
    uint32_t material_sort; /* 0 = off, 1 = sort hits by material inside each CTA tile, 2 = auto: on when the scene
                               holds more than one BSDF kind (measured: 1.57x on the shade kernel of a 7-kind scene,
                               a few % slower on a single-kind scene)                         */
    uint32_t sample_offset; /* index of the first sample: pass p of an averaging wrapper (avg.rs, equal_time.rs)
                               renders samples [p*spp, (p+1)*spp) so that passes never repeat a stream        */
} rl_render_opts;

typedef struct rl_stats {
    uint64_t samples;        /* W*H*spp rendered by this rank                                  */
    uint64_t segments;       /* closest-hit Acceleration::trace calls (edge.rs:90, direct.rs:33,144) */
    uint64_t shadow_rays;    /* Acceleration::visible calls (emitters.rs:125, direct.rs:75)    */
    uint64_t shadow_visible; /* of which unoccluded                                            */
    uint64_t hits;           /* closest-hit calls that found a surface                         */
    uint64_t max_depth_seen; /* deepest wavefront iteration                                    */
    uint64_t kernel_launches;
    double ms_total;  /* CUDA-event time of the whole rl_render device work                    */
    double ms_raygen, ms_trace, ms_shade, ms_shadow, ms_accum; /* filled when profiling is on   */
    double ms_h2d, ms_d2h, ms_reduce;
    double ms_tail;           /* k_tail launches (profiling on)                                 */
    uint64_t shadow_traced;   /* shadow segments actually traced (valid light samples with a non-zero contribution) */
    uint64_t launches_trace, launches_shade; /* launches behind ms_trace / ms_shade (profiling on) */
} rl_stats;

/* ---- context -------------------------------------------------------------------------------- */
/* Replaces rayon pool creation, integrators/mod.rs:452-459.  device = CUDA ordinal.
 * nranks/rank partition the image plane; nccl_unique_id (128 bytes from
 * rl_nccl_unique_id on rank 0) may be NULL when nranks == 1 or when the caller reduces the
 * framebuffer itself (see rl_render_device). */
int rl_create(rl_ctx **out, int device, int nranks, int rank, const void *nccl_unique_id);
void rl_destroy(rl_ctx *ctx);
int rl_nccl_unique_id(void *out_128_bytes);
const char *rl_last_error(const rl_ctx *ctx); /* ctx may be NULL: last error of rl_create */
int rl_abi_version(void);
/* Page-locked host memory for out_rgb (optional: rl_render accepts any host pointer; a pinned buffer lets the frame's
 * device-to-host copy run at PCIe/C2C line rate without a staging copy).  The reference returns its BufferCollection by
 * value (integrators/mod.rs:219-228); here the caller owns the buffer.  NULL on failure; rl_host_free(NULL) is a no-op. */
void *rl_host_alloc(size_t bytes);
void rl_host_free(void *p);
/* Per-launch CUDA-event timing (no synchronisation: events are read when the frame is complete).  0 = off; 1 = trace and
 * shadow kernels launched separately, so that ms_trace / ms_shadow are per stage; 2 = exactly the kernels of an untimed frame
 * (ms_trace then holds the fused closest-hit + shadow launches). */
int rl_set_profiling(rl_ctx *ctx, int on);

/* ---- scene ----------------------------------------------------------------------------------- */
/* Replaces BVHAccel::new (accel.rs:201-240) + Scene::build_emitters (scene.rs:53-123):
 * uploads geometry, builds the LBVH on the device, builds the flux-weighted emitter CDF. */
int rl_scene_create(rl_ctx *ctx, const rl_scene_desc *desc, rl_scene **out);
void rl_scene_destroy(rl_ctx *ctx, rl_scene *scene);

typedef struct rl_bvh_info {
    uint32_t ntris, nnodes, nleaves, max_depth;
    float root_min[3], root_max[3]; /* == BVHAccel nodes[0].aabb (accel.rs:204-217)            */
    uint32_t smem_resident;         /* 1 if BVH+triangles are staged in shared memory          */
    /* group table of small scenes (incoherent rays scan it instead of walking the tree): 0 groups = absent */
    uint32_t flat_groups, flat_pairs, flat_singles;
    float flat_delta;               /* largest plane mismatch inside a triangle pair           */
} rl_bvh_info;
int rl_scene_bvh_info(rl_ctx *ctx, const rl_scene *scene, rl_bvh_info *out);

/* ---- the hot path ---------------------------------------------------------------------------- */
/* Replaces Integrator::compute -> compute_mc (integrators/mod.rs:403-450).  out_rgb is a HOST
 * buffer of width*height*3 floats holding the mean radiance (1/spp already applied,
 * mod.rs:436), or NULL to leave the result on the device.  With nranks > 1 every rank renders
 * its tiles and one ncclReduce(sum) leaves the full image on rank 0 (out_rgb is written on
 * rank 0 only). */
int rl_render(rl_ctx *ctx, rl_scene *scene, const rl_integrator_desc *integrator,
              const rl_render_opts *opts, float *out_rgb, rl_stats *stats);
/* Same, but the (un-reduced, this rank's tiles only, zero elsewhere) mean image is written to
 * a DEVICE buffer of width*height*3 floats owned by the caller. */
int rl_render_device(rl_ctx *ctx, rl_scene *scene, const rl_integrator_desc *integrator,
                     const rl_render_opts *opts, float *out_rgb_device, rl_stats *stats);

/* Acceleration::trace (accel.rs:292-315) for n rays given as HOST arrays o[3n], d[3n]
 * (tnear = 1e-4, tfar = f32::MAX as Ray::new, structure.rs:705-715).  Outputs (HOST):
 * prim[n] = global triangle index (mesh-major) or 0xFFFFFFFF on miss, tuv[3n] = (t,u,v). */
int rl_trace(rl_ctx *ctx, rl_scene *scene, size_t n, const float *o, const float *d,
             uint32_t *prim, float *tuv);
/* Acceleration::visible (accel.rs:316-343) for n segments p0[3n] -> p1[3n]; out[n] = 0/1. */
int rl_visible(rl_ctx *ctx, rl_scene *scene, size_t n, const float *p0, const float *p1,
               uint8_t *out);
/* Camera::generate (camera.rs:81-91) + trace for the pixel-centre grid uv = (x+.5, y+.5):
 * prim[W*H] and tuv[3*W*H] (tuv may be NULL). */
int rl_primary_hits(rl_ctx *ctx, rl_scene *scene, uint32_t *prim, float *tuv);

/* Device pointers of the last batch's SoA queues are private; this returns sizes only. */
typedef struct rl_layout_info {
    uint32_t ray_bytes, hit_bytes, state_bytes, shadow_bytes, accum_bytes;
    uint64_t max_paths_in_flight;
} rl_layout_info;
int rl_layout(rl_ctx *ctx, rl_layout_info *out);

#ifdef __cplusplus
}
#endif
#endif /* RL_B200_H */
