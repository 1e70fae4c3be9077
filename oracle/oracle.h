/*
 * oracle.h -- C entry points of the CPU oracle (TEST INFRASTRUCTURE, NOT PRODUCT).
 *
 * The oracle is a C++ restatement of rustlight@864df34's `path` / `direct` hot path
 * (see oracle.cpp for the file:line map).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product library
 * (librl_b200.so) never links, loads or calls anything in this directory.
 *
 * PARITY UNPINNED: the reference ships no unit tests, golden vectors or known-answer tests
 * for this path and cannot be compiled here (no Rust toolchain),
 * so this restatement is pinned only by the analytic/self-consistency tests in tests/.
 */
#ifndef RL_ORACLE_H
#define RL_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#include "rl_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_scene orc_scene;

/* math_mode: which transcendental functions the sampling code uses. */
enum { ORC_MATH_LIBM = 0, /* glibc sinf/cosf/powf == what Rust's f32::sin/cos/powf call on Linux */
       ORC_MATH_SPEC = 1  /* the f64-polynomial sin/cos/pow defined in DESIGN.md, shared bit-for-bit with the GPU */ };
/* accel_mode */
enum { ORC_ACCEL_BVH = 0,  /* BVHAccel, src/accel.rs:101-344 (what the reference runs)        */
       ORC_ACCEL_NAIVE = 1 /* NaiveAcceleration trace, src/accel.rs:22-51, + BVHAccel's root test */ };
/* estimator */
enum { ORC_EST_GRAPH = 0,  /* build the path graph, then TechniquePathTracing::evaluate (path.rs:113-185) */
       ORC_EST_STREAM = 1  /* algebraically identical forward accumulation in the GPU's operation order */ };
/* mode-A seeding of SmallRng::seed_from_u64 (rand 0.8.5 source is not available offline) */
enum { ORC_SEED_PCG32 = 0, /* rand_core 0.6 default seed_from_u64 (believed to be what SmallRng 0.8.5 uses) */
       ORC_SEED_SPLITMIX64 = 1 /* Xoshiro256PlusPlus's own seed_from_u64 override */ };

typedef struct orc_config {
    uint32_t math_mode, accel_mode, estimator, seeding;
    uint32_t nthreads; /* 0 = all hardware threads */
    uint32_t rank, nranks; /* image-tile partition identical to the GPU's; nranks<=1 = whole image */
    uint32_t sample_offset; /* mode B: index of the first sample (rl_render_opts.sample_offset) */
} orc_config;

typedef struct orc_stats {
    uint64_t samples, segments, shadow_rays, shadow_visible, hits, max_depth_seen;
    uint64_t nee_added; /* ORC_EST_STREAM only: light-sampling contributions actually added (== GPU shadow_visible) */
    double seconds; /* wall clock of the block loop + merge == "Elapsed Integrator" region, mod.rs:323-334 */
    uint32_t threads_used;
} orc_stats;

orc_scene *orc_scene_create(const rl_scene_desc *desc, char *err, size_t errlen);
void orc_scene_destroy(orc_scene *s);

/* compute_mc (integrators/mod.rs:403-450) for IntegratorPathTracing / IntegratorDirect. */
int orc_render(const orc_scene *s, const rl_integrator_desc *integ, uint32_t spp, uint64_t seed,
               uint32_t sampler_mode, const orc_config *cfg, float *out_rgb, orc_stats *stats);

/* Acceleration::trace / visible for batches (same conventions as rl_trace / rl_visible);
 * extra outputs may be NULL: p[3n], n_g[3n], n_s[3n], wi[3n] from fill_intersection. */
int orc_trace(const orc_scene *s, uint32_t accel_mode, size_t n, const float *o, const float *d,
              uint32_t *prim, float *tuv, float *p, float *n_g, float *n_s, float *wi);
int orc_visible(const orc_scene *s, uint32_t accel_mode, size_t n, const float *p0, const float *p1, uint8_t *out);
int orc_primary_hits(const orc_scene *s, uint32_t accel_mode, uint32_t *prim, float *tuv);
void orc_bvh_info(const orc_scene *s, uint32_t *nnodes, uint32_t *nprims, float root_min[3], float root_max[3]);

/* ---- unit-level entry points for the known-answer tests ------------------------------------ */
/* Mesh::intersection_tri (geometry.rs:358-410) on an explicit triangle; t_io is its.t in/out. */
int orc_intersect_tri(const float v0[3], const float v1[3], const float v2[3], const float o[3], const float d[3],
                      float *t_io, float *u, float *v, float p[3], float n[3]);
/* AABB::intersect (structure.rs:849-869): returns 1 and *t on hit. */
int orc_aabb_intersect(const float pmin[3], const float pmax[3], const float o[3], const float d[3], float tnear,
                       float tfar, float *t);
void orc_frame(const float n[3], float out9[9]); /* Frame::new, math.rs:360-371: x,y,z columns */
void orc_cosine_sample_hemisphere(uint32_t math_mode, float u0, float u1, float out[3]); /* math.rs:37-65 */
void orc_uniform_sample_triangle(float u0, float u1, float out[2]);                       /* math.rs:388-394 */
/* Distribution1DConstruct::normalize (math.rs:418-441): cdf has n+1 entries; returns func_int. */
float orc_dist1d_normalize(const float *elements, uint32_t n, float *cdf);
uint32_t orc_dist1d_sample_discrete(const float *cdf, uint32_t n_plus_1, float v); /* math.rs:447-457 */
float orc_mis_weight(float pdf_a, float pdf_b);                                    /* integrators/mod.rs:462-478 */
/* BSDF::{sample,pdf,eval} for an rl_material.  sample: returns 1 if Some; out = weight[3], d[3], pdf. */
int orc_bsdf_sample(uint32_t math_mode, const rl_material *m, const float wi[3], float s0, float s1,
                    float weight[3], float d[3], float *pdf);
float orc_bsdf_pdf(uint32_t math_mode, const rl_material *m, const float wi[3], const float wo[3]);
int orc_bsdf_flags(const rl_material *m); /* bit 0 is_twosided, bit 1 is_smooth */
void orc_bsdf_eval(uint32_t math_mode, const rl_material *m, const float wi[3], const float wo[3], float out[3]);
/* EmitterSampler::sample_light (emitter.rs:1604-1620).  Outputs: p[3], n[3], d[3], weight[3], pdf; returns emitter mesh index. */
int orc_sample_light(const orc_scene *s, const float x[3], float r_sel, float r, float u0, float u1, float p[3],
                     float n[3], float d[3], float weight[3], float *pdf);
/* EmitterSampler::direct_pdf (emitter.rs:1566-1575) for a point p with normal n on mesh `mesh`, seen from o along dir. */
float orc_direct_pdf(const orc_scene *s, uint32_t mesh, const float o[3], const float p[3], const float n[3],
                     const float dir[3]);
/* Camera::new (camera.rs:31-67) and Camera::generate (camera.rs:81-91). */
int orc_camera_new(uint32_t w, uint32_t h, int fov_axis, float fov_deg, const float to_world[16], int flip,
                   float sample_to_camera[16]);
void orc_camera_generate(const orc_scene *s, float px, float py, float o[3], float d[3]);
/* Samplers.  mode A: master seed -> per-block clone in x-major order -> first n draws of block `block`. */
void orc_sampler_block_stream(uint64_t seed, uint32_t seeding, uint32_t block, uint32_t n, float *out);
/* mode B: first n draws of (seed, pixel, sample). */
void orc_sampler_counter(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t n, float *out);
uint64_t orc_xoshiro_next_u64(uint64_t state[4]); /* xoshiro256++ step (KAT against the published algorithm) */
void orc_spec_sincos(float x, float *s, float *c);
/* LightSamplerATS of a scene created with use_ats (emitter.rs:1319-1399): sample(r) -> global triangle index + pdf; pdf of a triangle.
 * has_n = 0 passes None for the normal.  -1 when the scene has no light tree. */
int orc_ats_sample(const orc_scene *s, float r, const float p[3], const float n[3], int has_n, uint32_t *prim, float *pdf);
int orc_ats_pdf(const orc_scene *s, uint32_t prim, const float p[3], const float n[3], int has_n, float *pdf);
float orc_spec_atan2(float y, float x);
float orc_spec_acos(float x);
/* EnvironmentLightColor of the scene's environment (emitter.rs:354-425): eval / pdf of a direction, sample_direction(uv) -> d, colour, pdf */
int orc_env_eval_pdf(const orc_scene *s, uint32_t math_mode, const float d[3], float rgb[3], float *pdf);
int orc_env_sample(const orc_scene *s, uint32_t math_mode, float u0, float u1, float d[3], float rgb[3], float *pdf);
float orc_spec_powf(float x, float y);
/* A single path sample with full trace of what happened (for debugging parity):
 * returns radiance, and writes up to cap (segments, shadow rays) counters. */
void orc_path_sample(const orc_scene *s, const rl_integrator_desc *integ, uint64_t seed, uint32_t px, uint32_t py,
                     uint32_t sample, const orc_config *cfg, float rgb[3], uint32_t *n_segments, uint32_t *n_shadow,
                     uint32_t *n_draws);

#ifdef __cplusplus
}
#endif
#endif
