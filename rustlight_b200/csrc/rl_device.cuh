// rl_device.cuh -- per-thread arithmetic of the wavefront path tracer.
//
// Everything here is a plain inline function over registers and read-only scene tables, so
// the same source is compiled (a) by nvcc for the kernels in rl_kernels.cu (-fmad=false:
// every f32/f64 operation is one IEEE-754 operation in source order) and (b) by g++ for the
// serial emulator in tests/emu (RL_HD expands to nothing) that lets the CPU test-suite check
// the device arithmetic against the oracle without a GPU.  It is NOT a CPU fallback: the
// product library only ever calls these from __global__ kernels.
//
// Operation order follows the reference so that discrete decisions (hit / miss, visible /
// occluded, Russian roulette, lobe choice) are bit-identical to the CPU restatement:
//   Mesh::intersection_tri           src/geometry.rs:358-410
//   AABB::intersect                  src/structure.rs:849-869   (root test only)
//   Intersection::fill_intersection  src/structure.rs:965-1059
//   Frame                            src/math.rs:357-384
//   concentric/cosine sampling       src/math.rs:37-65
//   BSDFDiffuse / BSDFPhong          src/bsdfs/diffuse.rs, src/bsdfs/phong.rs
//   Mesh::sample_tri/sample, direct_sample, direct_pdf   src/geometry.rs:261-348, src/emitter.rs:571-688
//   EmitterSampler::sample_light/direct_pdf              src/emitter.rs:1566-1647
//   Camera::generate                 src/camera.rs:81-91
//   Color ops                        src/structure.rs:106-380
#pragma once
#include <stdint.h>
#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#define RL_HD __host__ __device__ __forceinline__
#define RL_HD_NOINLINE __host__ __device__ __noinline__ // rare slow paths: keep their registers and local arrays out of the callers
#else
#define RL_HD inline
#define RL_HD_NOINLINE inline
struct float4 {
    float x, y, z, w;
};
struct float2 {
    float x, y;
};
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
#endif

namespace rl {

#define RL_EPSILON 0.0001f
#define RL_F32_MAX 3.402823466e+38f
#define RL_PI 3.14159265358979323846264338327950288f
#define RL_FRAC_PI_2 1.57079632679489661923132169163975144f
#define RL_FRAC_PI_4 0.785398163397448309615660845819875721f
#define RL_FRAC_1_PI 0.318309886183790671537767526745028724f
#define RL_MISS 0xFFFFFFFFu
#define RL_PRIM_NEEDLE 0x40000000u // flag in the prim word of a trav[] record (rl_build.cuh: tri_setup)
#define RL_PRIM_MASK 0x01ffffffu   // triangle index (< 2^25)

RL_HD uint32_t f2u(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
#endif
}
RL_HD float u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}
RL_HD uint64_t d2u(double d) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t u;
    memcpy(&u, &d, 8);
    return u;
#endif
}
RL_HD double u2d(uint64_t u) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double d;
    memcpy(&d, &u, 8);
    return d;
#endif
}
RL_HD bool finite_f(float x) { return (f2u(x) & 0x7f800000u) != 0x7f800000u; }

// ---- vectors (cgmath op order: dot = (xx+yy)+zz, normalize = v * (1/|v|)) ------------------
struct V3 {
    float x, y, z;
};
RL_HD V3 v3(float x, float y, float z) { return V3{x, y, z}; }
RL_HD V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
RL_HD V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
RL_HD V3 operator-(V3 a) { return V3{-a.x, -a.y, -a.z}; }
RL_HD V3 operator*(V3 a, float s) { return V3{a.x * s, a.y * s, a.z * s}; }
RL_HD V3 operator*(float s, V3 a) { return V3{s * a.x, s * a.y, s * a.z}; }
RL_HD V3 operator/(V3 a, float s) { return V3{a.x / s, a.y / s, a.z / s}; }
RL_HD float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
RL_HD V3 cross(V3 a, V3 b) { return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
RL_HD float magnitude(V3 a) { return sqrtf(dot(a, a)); }
RL_HD V3 normalize(V3 a) { return a * (1.0f / magnitude(a)); }
RL_HD V3 xyz(float4 f) { return V3{f.x, f.y, f.z}; }

// ---- Color (structure.rs:106-380) -----------------------------------------------------------
struct Col {
    float r, g, b;
};
RL_HD Col col(float r, float g, float b) { return Col{r, g, b}; }
RL_HD bool is_zero(Col c) { return c.r == 0.0f && c.g == 0.0f && c.b == 0.0f; }
RL_HD float channel_max(Col c) { return fmaxf(c.r, fmaxf(c.g, c.b)); }
RL_HD Col operator*(Col a, Col b) { return Col{a.r * b.r, a.g * b.g, a.b * b.b}; }
RL_HD Col operator+(Col a, Col b) { return Col{a.r + b.r, a.g + b.g, a.b + b.b}; }
// Color * f32 returns zero when the scalar is not finite (structure.rs:278-292)
RL_HD Col mul_checked(Col a, float s) { return finite_f(s) ? Col{a.r * s, a.g * s, a.b * s} : Col{0.0f, 0.0f, 0.0f}; }
// f32 * Color has no such check (structure.rs:294-303)
RL_HD Col mul_plain(float s, Col a) { return Col{a.r * s, a.g * s, a.b * s}; }
// Color / f32 returns zero on a zero or non-finite divisor (structure.rs:249-265)
RL_HD Col div_checked(Col a, float s) { return (s == 0.0f || !finite_f(s)) ? Col{0.0f, 0.0f, 0.0f} : Col{a.r / s, a.g / s, a.b / s}; }
RL_HD Col xyz_col(float4 f) { return Col{f.x, f.y, f.z}; }

// ---- "spec" transcendental functions: DESIGN.md §math.  f64 polynomials; the oracle carries
// an independently typed copy with the same operation sequence (oracle.cpp spec_*). ----------
// sin and cos in f32: Cody-Waite reduction by pi/2 in three exact pieces (valid far beyond the |x| <= 2 pi the callers
// pass), then the Cephes sinf / cosf kernels on |y| <= pi/4, every step one fmaf / mul (identical on the device, in the
// emulator and in the oracle).  Max error ~1 ulp: the LIBM-vs-SPEC gate (tests: test_math_modes_agree) stays below 1e-5.
RL_HD void spec_sincos(float x, float *s, float *c) {
    const float fn = rintf(x * 0.636619772f);
    const int n = (int)fn;
    float y = fmaf(fn, -1.5703125f, x);
    y = fmaf(fn, -4.837512969970703125e-4f, y);
    y = fmaf(fn, -7.54978995489188e-8f, y);
    const float y2 = y * y;
    float ps = -1.9515295891e-4f;
    ps = fmaf(ps, y2, 8.3321608736e-3f);
    ps = fmaf(ps, y2, -1.6666654611e-1f);
    const float sy = fmaf(y * y2, ps, y);
    float pc = 2.443315711809948e-5f;
    pc = fmaf(pc, y2, -1.388731625493765e-3f);
    pc = fmaf(pc, y2, 4.166664568298827e-2f);
    const float cy = fmaf(y2 * y2, pc, fmaf(-0.5f, y2, 1.0f));
    const int q = n & 3;
    float rs, rc;
    if (q == 0) { rs = sy; rc = cy; }
    else if (q == 1) { rs = cy; rc = -sy; }
    else if (q == 2) { rs = -sy; rc = -cy; }
    else { rs = -cy; rc = sy; }
    *s = rs;
    *c = rc;
}
// atan2 and acos in f32 (EnvironmentLightColor::Texture: to_spherical_coordinates, emitter.rs:320-338): the Cephes atanf / asinf
// kernels, every step one fmaf / mul / div / sqrt, identical on the device, in the emulator and in the oracle.  ~2 ulp.
RL_HD float spec_atanf_pos(float x) { // x >= 0
    float y0, z;
    if (x > 2.414213562373095f) {
        y0 = 1.5707963267948966f;
        z = -(1.0f / x);
    } else if (x > 0.4142135623730950f) {
        y0 = 0.7853981633974483f;
        z = (x - 1.0f) / (x + 1.0f);
    } else {
        y0 = 0.0f;
        z = x;
    }
    const float z2 = z * z;
    float p = 8.05374449538e-2f;
    p = fmaf(p, z2, -1.38776856032e-1f);
    p = fmaf(p, z2, 1.99777106478e-1f);
    p = fmaf(p, z2, -3.33329491539e-1f);
    return y0 + fmaf(p * z2, z, z);
}
RL_HD float spec_atan2f(float y, float x) { // result in (-pi, pi]; (0, 0) -> 0
    if (x != x || y != y) return x + y;
    if (x == 0.0f) return y > 0.0f ? 1.5707963267948966f : (y < 0.0f ? -1.5707963267948966f : 0.0f);
    const float q = y / x;
    const float a = spec_atanf_pos(fabsf(q));
    const float at = q < 0.0f ? -a : a;
    if (x > 0.0f) return at;
    return (f2u(y) >> 31) ? at - 3.14159265358979323846f : at + 3.14159265358979323846f; // sign bit: atan2(-0, x < 0) = -pi
}
RL_HD float spec_asinf_pos(float a) { // 0 <= a <= 1
    if (a < 1e-4f) return a;
    const bool big = a > 0.5f;
    float z, x;
    if (big) {
        z = 0.5f * (1.0f - a);
        x = sqrtf(z);
    } else {
        x = a;
        z = x * x;
    }
    float p = 4.2163199048e-2f;
    p = fmaf(p, z, 2.4181311049e-2f);
    p = fmaf(p, z, 4.5470025998e-2f);
    p = fmaf(p, z, 7.4953002686e-2f);
    p = fmaf(p, z, 1.6666752422e-1f);
    float r = fmaf(p * z, x, x);
    if (big) r = 1.5707963267948966f - (r + r);
    return r;
}
RL_HD float spec_acosf(float x) { // |x| <= 1 (callers clamp); NaN in, NaN out
    if (x != x) return x;
    if (x < -0.5f) return 3.14159265358979323846f - 2.0f * spec_asinf_pos(sqrtf(0.5f * (1.0f + x)));
    if (x > 0.5f) return 2.0f * spec_asinf_pos(sqrtf(0.5f * (1.0f - x)));
    const float a = spec_asinf_pos(fabsf(x));
    return 1.5707963267948966f - (x < 0.0f ? -a : a);
}
RL_HD double spec_log2(double x) {
    uint64_t bits = d2u(x);
    int e = (int)((bits >> 52) & 0x7ff) - 1023;
    double m = u2d((bits & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL);
    if (m > 1.4142135623730951) {
        m = m * 0.5;
        e = e + 1;
    }
    double f = (m - 1.0) / (m + 1.0);
    double f2 = f * f;
    double p = 1.0 / 21.0;
    p = p * f2 + 1.0 / 19.0;
    p = p * f2 + 1.0 / 17.0;
    p = p * f2 + 1.0 / 15.0;
    p = p * f2 + 1.0 / 13.0;
    p = p * f2 + 1.0 / 11.0;
    p = p * f2 + 1.0 / 9.0;
    p = p * f2 + 1.0 / 7.0;
    p = p * f2 + 1.0 / 5.0;
    p = p * f2 + 1.0 / 3.0;
    double ln_m = 2.0 * (f + f * (f2 * p));
    return (double)e + ln_m * 1.4426950408889634074;
}
RL_HD double spec_exp2(double t) {
    if (t < -1000.0) return 0.0;
    if (t > 1000.0) return u2d(0x7ff0000000000000ULL);
    double k = floor(t + 0.5);
    double r = (t - k) * 0.69314718055994530942;
    double p = 1.0 / 6227020800.0;
    p = p * r + 1.0 / 479001600.0;
    p = p * r + 1.0 / 39916800.0;
    p = p * r + 1.0 / 3628800.0;
    p = p * r + 1.0 / 362880.0;
    p = p * r + 1.0 / 40320.0;
    p = p * r + 1.0 / 5040.0;
    p = p * r + 1.0 / 720.0;
    p = p * r + 1.0 / 120.0;
    p = p * r + 1.0 / 24.0;
    p = p * r + 1.0 / 6.0;
    p = p * r + 0.5;
    p = p * r + 1.0;
    p = p * r + 1.0;
    int ki = (int)k;
    int k1 = ki / 2, k2 = ki - k1;
    double s1 = u2d((uint64_t)(k1 + 1023) << 52), s2 = u2d((uint64_t)(k2 + 1023) << 52);
    return (p * s1) * s2;
}
RL_HD float spec_powf(float x, float y) {
    if (y == 0.0f) return 1.0f;
    if (x == 0.0f) return y > 0.0f ? 0.0f : u2f(0x7f800000u);
    if (x == 1.0f) return 1.0f;
    if (!(x > 0.0f)) return u2f(0x7fc00000u);
    return (float)spec_exp2((double)y * spec_log2((double)x));
}

// spec_powf(x, y) with log2(x) supplied by the caller (same bits: the same log2 value enters the same product), so that
// several powers of one base share the logarithm
RL_HD float spec_powf_pre(float x, float y, double log2_x) {
    if (y == 0.0f) return 1.0f;
    if (x == 0.0f) return y > 0.0f ? 0.0f : u2f(0x7f800000u);
    if (x == 1.0f) return 1.0f;
    if (!(x > 0.0f)) return u2f(0x7fc00000u);
    return (float)spec_exp2((double)y * log2_x);
}

// e^x and ln x (Beckmann distribution only), same construction: f64 kernels, one rounding to f32
RL_HD float spec_expf(float x) { return (float)spec_exp2((double)x * 1.4426950408889634074); }
RL_HD float spec_logf(float x) {
    if (x == 0.0f) return -u2f(0x7f800000u);
    if (!(x > 0.0f)) return u2f(0x7fc00000u);
    if (!finite_f(x)) return x;
    return (float)(spec_log2((double)x) * 0.69314718055994530942);
}

// ---- counter-based sampler (mode B) -----------------------------------------
RL_HD uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
RL_HD uint64_t seed_hash(uint64_t seed) { return mix64(seed + 0x9e3779b97f4a7c15ULL); }
// Draw number n (1-based) of a pixel sample is one 24-bit half of hash number h = (n + 1) / 2: bits 40..63 for odd n,
// bits 16..39 for even n -- one SplitMix64 finaliser per TWO draws.  The sampler keeps the last hash.
struct Sampler {
    uint64_t key, z;
    uint32_t n, zh;
    RL_HD float next() {
        n++;
        const uint32_t h = (n + 1u) >> 1;
        if (h != zh) {
            z = mix64(key + (uint64_t)h * 0x9e3779b97f4a7c15ULL);
            zh = h;
        }
        const uint32_t bits = (n & 1u) ? (uint32_t)(z >> 40) : ((uint32_t)(z >> 16) & 0xffffffu);
        return (float)bits * (1.0f / 16777216.0f);
    }
};
// The same stream read in (even, odd) slot order from a start n0 that is only known at run time, with every hash at a fixed place in
// the code: no "is the hash still valid" test and no branch per draw (k_shade: the generic next() was 7 % of the kernel's
// instructions).  With s = n0 & 1, slot j of this reader is half j + s of the hash sequence that starts at hash number n0 / 2 + 1:
// an even slot reads the current hash (upper half when s = 0, lower when s = 1), an odd slot computes the next hash and reads its
// upper half (s = 1) or the current hash's lower half (s = 0).  Calls must alternate even(), odd(), even(), ...; after a draw that is
// taken conditionally the caller re-begins at the new count (one more hash).
#ifndef RL_PAIR_SAMPLER
#define RL_PAIR_SAMPLER 1
#endif
struct PairSampler {
    uint64_t key, zc;
    uint32_t h, n;
    bool s;
    RL_HD static uint32_t hi24(uint64_t z) { return (uint32_t)(z >> 40); }
    RL_HD static uint32_t lo24(uint64_t z) { return (uint32_t)(z >> 16) & 0xffffffu; }
    RL_HD static float unit(uint32_t bits) { return (float)bits * (1.0f / 16777216.0f); }
    RL_HD void begin(uint32_t n0) {
        n = n0;
        s = (n0 & 1u) != 0u;
        h = (n0 >> 1) + 1u;
        zc = mix64(key + (uint64_t)h * 0x9e3779b97f4a7c15ULL);
    }
    RL_HD float even() {
        n++;
        return unit(s ? lo24(zc) : hi24(zc));
    }
    RL_HD float odd() {
        n++;
        h++;
        const uint64_t zn = mix64(key + (uint64_t)h * 0x9e3779b97f4a7c15ULL);
        const uint32_t bits = s ? hi24(zn) : lo24(zc);
        zc = zn;
        return unit(bits);
    }
};
RL_HD PairSampler make_pair_sampler(uint64_t seed_h, uint32_t pixel, uint32_t sample, uint32_t n) {
    PairSampler p;
    p.key = mix64(seed_h ^ (((uint64_t)pixel << 32) | (uint64_t)sample));
    p.begin(n);
    return p;
}
RL_HD Sampler make_sampler(uint64_t seed_h, uint32_t pixel, uint32_t sample, uint32_t n) {
    Sampler s;
    s.key = mix64(seed_h ^ (((uint64_t)pixel << 32) | (uint64_t)sample));
    s.n = n;
    s.z = 0ull, s.zh = 0u; // no hash yet (hash numbers start at 1)
    return s;
}

// ---- Frame (math.rs:357-384) ------------------------------------------------------------------
struct Frame {
    V3 x, y, z;
};
RL_HD Frame make_frame(V3 n) {
    float sign = copysignf(1.0f, n.z);
    float a = -1.0f / (sign + n.z);
    float b = n.x * n.y * a;
    Frame f;
    f.x = V3{1.0f + sign * n.x * n.x * a, sign * b, -sign * n.x};
    f.y = V3{b, sign + n.y * n.y * a, -n.y};
    f.z = n;
    return f;
}
RL_HD V3 to_world(const Frame &f, V3 v) { return f.x * v.x + f.y * v.y + f.z * v.z; }
RL_HD V3 to_local(const Frame &f, V3 v) { return V3{dot(v, f.x), dot(v, f.y), dot(v, f.z)}; }

// ---- sampling (math.rs:37-65, 388-394) -------------------------------------------------------
RL_HD V3 cosine_sample_hemisphere(float ux, float uy) {
    float ox = ux * 2.0f - 1.0f, oy = uy * 2.0f - 1.0f;
    float dx, dy;
    if (ox == 0.0f && oy == 0.0f) {
        dx = 0.0f;
        dy = 0.0f;
    } else {
        float theta, r;
        if (fabsf(ox) > fabsf(oy)) {
            r = ox;
            theta = RL_FRAC_PI_4 * (oy / ox);
        } else {
            r = oy;
            theta = RL_FRAC_PI_2 - RL_FRAC_PI_4 * (ox / oy);
        }
        float s, c;
        spec_sincos(theta, &s, &c);
        dx = c * r;
        dy = s * r;
    }
    float z = sqrtf(fmaxf(0.0f, 1.0f - dx * dx - dy * dy));
    return V3{dx, dy, z};
}

// ---- scene tables ------------------------------------------------------------------------------
// All tables are arrays of float4 so that every access is one 128-bit load.
//   trav[4*s+0..3]   (Morton order s): {v0.xyz, det} {e1.xyz, prim} {e2.xyz, -} {n_geo.xyz, -}
//   nodes[4*j+0..3]  wide LBVH node:   {lo0.xyz, hi0.x} {hi0.yz, lo1.xy} {lo1.z, hi1.xyz} {child0, child1, -, -}
//   shade[4*p+0..3]  (original order p): {n_geo.xyz, mesh} {n0.xyz, flags: bit 0 normals, bit 1 uv} {n1.xyz, -} {n2.xyz, -}
//   uvs[3*p+0..2]    per-corner texture coordinates (only when some mesh has uv)
//   tex[4*t+0..3]    {color0, kind} {color1, line_width} {offset.xy, scale.xy} {width, height, texel offset, -};  texels[]
//   verts[3*p+0..2]  (original order p): {v.xyz, -}
//   mats[5*m+0..4]   {kd.rgb | metal eta | glass kt, kind} {ks.rgb, phong exponent | microfacet alpha} {Le.rgb, is_light}
//                    {phong weight_specular | glass eta, 1/area, pdf_sel, microfacet} {metal k.rgb, glass 1/eta}
//   emit_info[2e..]  mesh light {mesh, first_prim, ntris, cdf_offset} {-};  point / directional light
//                    {0xfffffff0 | rl_light_kind, intensity.rgb} {position | direction, bounding-sphere radius}
struct SceneView {
    const float4 *trav;
    const float4 *nodes;
    const float4 *shade;
    const float4 *verts;
    const float4 *mats;
    const float4 *emit_info;
    const float2 *uvs;    // nullptr when no mesh has uv
    const float4 *tex;    // textures of the kd slot (BSDFColor::{Bitmap, Checkerbord, Grid})
    uint32_t emit_var;    // some mesh light has EmissionType::HSV / Texture (mesh_emit)
    const float4 *texels;
    const float *emit_cdf; // n_emitters+1
    const float *area_cdf; // concatenated per-emitter triangle-area cdfs (ntris+1 each)
    uint32_t ntris, n_emitters;
    // flat group table (rl_build.cuh: build_flat_table) for scenes of a few dozen triangles: n_groups == 0 when absent
    const float4 *flat;
    uint32_t n_groups;
    uint32_t flat_valid_a, flat_valid_b; // quad bits whose A / B triangle is real (bit order of flat_scan)
    float flat_delta;       // largest plane mismatch inside a triangle pair (world units), added to the t margin
    // the reference's own BVH (rl_refbvh_host.hpp), walked only by rays whose two nearest accepted hits (nearly) tie: nullptr = absent
    const float4 *ref_nodes;   // 2 per node: {p_min, info} {p_max, count}
    const uint32_t *ref_prims; // leaf contents as Morton slots (indices into trav[])
    const uint32_t *ref_up;    // [node] parent; [ref_n_nodes + Morton slot] the leaf that holds the triangle
    uint32_t ref_n_nodes;
    int root_ref; // inner node 0, or a leaf reference when the whole scene is one leaf
    uint32_t wide4; // nodes[] holds 4-wide nodes (seven float4 each, rl_wide_host.hpp) instead of the 64-byte two-child nodes
    // BVHAccel nodes[0].aabb (union of compute_aabb_tri boxes), for the reference's root test
    V3 root_min, root_max;
    float abs_max; // max |coordinate| of the scene, scales the conservative-culling epsilon
    // EnvironmentLight with a constant colour (emitter.rs:428-568): 0 when absent
    uint32_t env_on;
    Col env_color;
    V3 bs_center;      // Scene.bsphere (scene.rs:54-60) ...
    float bs_radius;   // ... radius x 1.1 (EnvironmentLight::preprocess)
    float env_pdf_sel; // probability of picking the environment in sample_light
    // EnvironmentLightColor::Texture (emitter.rs:300-427): env_w == 0 when the environment is constant.  env_dist = the Distribution2D
    // (math.rs:489-532) as one float array: marginal cdf [env_h + 1], conditional cdfs [env_h][env_w + 1], conditional funcs [env_h][env_w]
    uint32_t env_w, env_h, env_texel_off; // image size and its offset in texels[]
    const float *env_dist;
    float env_func_int; // marginal.func_int
    // LightSamplerATS (emitter.rs:1130-1400; rl_ats_host.hpp): four float4 per node, nullptr without `-x ats`
    const float4 *ats_nodes;
    const uint32_t *ats_leaf_of_prim; // [global triangle] -> leaf node (0xffffffff: not a light)
    uint32_t ats_root;
    // camera
    float s2c[16], c2w[16];
    V3 cam_pos;
    float img_w, img_h;
};

// ---- AABB::intersect (structure.rs:849-869), used verbatim for the root box -------------------
// `inv` must be the IEEE quotients 1.0f/d.{x,y,z} (the reference computes them inside the loop;
// they are hoisted so that the LBVH slab test can reuse them).
RL_HD bool aabb_intersect_ref(V3 pmin, V3 pmax, V3 o, V3 inv, float tnear, float tfar) {
    float t_max = tfar, t_min = tnear;
    {
        float t0 = (pmin.x - o.x) * inv.x, t1 = (pmax.x - o.x) * inv.x;
        if (inv.x < 0.0f) { float t = t0; t0 = t1; t1 = t; }
        t_min = t0 > t_min ? t0 : t_min;
        t_max = t1 < t_max ? t1 : t_max;
        if (t_max <= t_min) return false;
    }
    {
        float t0 = (pmin.y - o.y) * inv.y, t1 = (pmax.y - o.y) * inv.y;
        if (inv.y < 0.0f) { float t = t0; t0 = t1; t1 = t; }
        t_min = t0 > t_min ? t0 : t_min;
        t_max = t1 < t_max ? t1 : t_max;
        if (t_max <= t_min) return false;
    }
    {
        float t0 = (pmin.z - o.z) * inv.z, t1 = (pmax.z - o.z) * inv.z;
        if (inv.z < 0.0f) { float t = t0; t0 = t1; t1 = t; }
        t_min = t0 > t_min ? t0 : t_min;
        t_max = t1 < t_max ? t1 : t_max;
        if (t_max <= t_min) return false;
    }
    return true;
}

// ---- Mesh::intersection_tri with the ray-independent terms (e1, e2, n_geo, det) precomputed ---
// Exactly the reference's accept/reject decision and (t,u,v) values, with the tests re-ordered
// (every test is a pure function of the inputs, so order cannot change the outcome): `t` is
// compared with the caller's bound first (the reference does it last, geometry.rs:398).
// `t_bound` semantics: closest hit passes best.t and accepts t <= t_bound here (the caller
// resolves the t == best.t tie by triangle index); shadow rays pass thr and the caller checks <.
RL_HD bool tri_test(float4 r0, float4 r1, float4 r2, float4 r3, V3 o, V3 d, float t_bound, float *t_out, float *u_out, float *v_out) {
    V3 v0 = xyz(r0), e1 = xyz(r1), e2 = xyz(r2), n_geo = xyz(r3);
    float det = r0.w;
    float denom = dot(d, n_geo);
    if (denom == 0.0f) return false;
    float t = -dot(o - v0, n_geo) / denom;
    if (t < 0.0f) return false;
    if (!(t <= t_bound) || !(t > 0.00001f)) return false;
    V3 p = o + t * d;
    V3 pv = p - v0;
    V3 u0 = cross(e1, pv);
    if (dot(u0, n_geo) < 0.0f) return false;
    V3 v0c = cross(pv, e2);
    if (dot(v0c, n_geo) < 0.0f) return false;
    float v = magnitude(u0) / det;
    float u = magnitude(v0c) / det;
    if (u < 0.0f || v < 0.0f || u > 1.0f || v > 1.0f) return false;
    if (!(u + v <= 1.0f)) return false;
    *t_out = t;
    *u_out = u;
    *v_out = v;
    return true;
}

// ---- exact ties: the reference's visit order ------------------------------------------------------
// Two accepted hits whose t differ by at most RL_TIE_WINDOW (relative) are "tied": which one the reference returns depends on
// the order in which BVHAccel::intersect visits its leaves and on its `d < its.t` culling (accel.rs:243-288), far beyond what
// a distance comparison can tell.  Hits further apart than the window are ordered identically by every correct traversal
// (a box that holds the nearer triangle is entered before the farther hit's t, rounding is ~1e-7 t).
#define RL_TIE_WINDOW 2e-5f
#ifndef RL_REF_ORDER
#define RL_REF_ORDER 1 // 0 compiles the reference-order slow path out (A/B hook: what does carrying it cost?)
#endif
#ifndef RL_FLAT_MARGIN_SCALE
#define RL_FLAT_MARGIN_SCALE 1.0f // test hook: tests/ shrink the margins to measure how much slack they carry
#endif
RL_HD float tie_bound(float t) { return t + t * RL_TIE_WINDOW; } // t >= 0; F32_MAX -> inf
RL_HD bool tie_window(float a, float b) { return fabsf(a - b) <= RL_TIE_WINDOW * fmaxf(a, b); }
// AABB::intersect returning the entry distance (structure.rs:849-869)
RL_HD bool aabb_entry_ref(V3 pmin, V3 pmax, V3 o, V3 inv, float tnear, float tfar, float *t_entry) {
    float t_max = tfar, t_min = tnear;
    {
        float t0 = (pmin.x - o.x) * inv.x, t1 = (pmax.x - o.x) * inv.x;
        if (inv.x < 0.0f) { float t = t0; t0 = t1; t1 = t; }
        t_min = t0 > t_min ? t0 : t_min;
        t_max = t1 < t_max ? t1 : t_max;
        if (t_max <= t_min) return false;
    }
    {
        float t0 = (pmin.y - o.y) * inv.y, t1 = (pmax.y - o.y) * inv.y;
        if (inv.y < 0.0f) { float t = t0; t0 = t1; t1 = t; }
        t_min = t0 > t_min ? t0 : t_min;
        t_max = t1 < t_max ? t1 : t_max;
        if (t_max <= t_min) return false;
    }
    {
        float t0 = (pmin.z - o.z) * inv.z, t1 = (pmax.z - o.z) * inv.z;
        if (inv.z < 0.0f) { float t = t0; t0 = t1; t1 = t; }
        t_min = t0 > t_min ? t0 : t_min;
        t_max = t1 < t_max ? t1 : t_max;
        if (t_max <= t_min) return false;
    }
    *t_entry = t_min;
    return true;
}
// Is the accepted hit (u, v) within the rounding-safe margin of the triangle's boundary?  The reference's boxes are the exact
// vertex bounds: a hit on a vertex or edge that lies on a box face can fail BVHAccel's slab test by one ulp (t_max <= t_min),
// so the reference MISSES it.  A hit further inside than rim_margin (in barycentrics; 2e-6 x the coordinate scale in space, ~10x the
// rounding of the slab test) is strictly inside every box on its root-to-leaf path and is found by the reference too.
RL_HD float rim_margin(float mn, float rs8) { return fmaf(mn, 0.25f * rs8, RL_FLAT_MARGIN_SCALE * 2e-6f); } // 2e-6 x coordinate scale in space
RL_HD bool hit_near_edge(float u, float v, float mn, float rs8) { return !(fminf(fminf(u, v), 1.0f - u - v) >= rim_margin(mn, rs8)); }
// The other ways BVHAccel can lose an interior hit, all of them properties of its boxes (geometry.rs:423-439: exact vertex bounds,
// flat extents padded by 1e-4) and of its slab test with tnear = 1e-4 (structure.rs:849-869):
//   * t <= 1e-4: a box the ray leaves before tnear fails `t_max <= t_min`;
//   * a box that is thin along two axes (RL_PRIM_NEEDLE): the chord of the ray inside it can be shorter than the rounding of the test;
//   * the 1e-4 padding itself is no longer large against that rounding, ~3 ulp of (2 |o| + t + scene extent).
// A hit that is none of these (and neither tied nor on the rim) lies inside every box on its root-to-leaf path with a chord far
// longer than the rounding, so the reference finds it: the device's answer IS the reference's.  Everything else is re-traced by
// ref_bvh_closest / ref_bvh_any, i.e. by the reference's own algorithm.
RL_HD bool hit_unsafe(uint32_t prim_word, float t, float omax, float abs_max) {
    const float R = 2e-7f * (2.0f * omax + t + abs_max);
    return (prim_word & RL_PRIM_NEEDLE) != 0u || !(t > 1.001e-4f) || !(8.0f * R <= RL_EPSILON);
}

// ---- LBVH traversal state -----------------------------------------------------------------------
// The reference's BVH shape is not reproduced.  Culling is conservative:
// node boxes are pre-inflated by 3e-5*abs_max at build time (rl_build.cuh), far more than the
// rounding of the fma-based slab test below, and rays whose origin lies far outside the scene
// add a per-ray term.  A culling test that is merely conservative cannot change the result,
// which is fixed by the exact triangle test + the tie rule: closest accepted triangle, lowest
// (mesh,tri) index on exact ties == the brute-force loop of NaiveAcceleration (accel.rs:22-51).
#ifndef RL_STACK_SIZE
#define RL_STACK_SIZE 64
#endif
#define RL_TRAV_DONE 0x7fffffff
#define RL_TRAV_EMPTY 0x7ffffffe // unused child slot of a 4-wide node (rl_wide_host.hpp)
#define RL_LEAF_MAX_CAP 64
struct int2v {
    int x, y;
};
// leaf reference: bit 31 | (count-1) << 25 | first   (first < 2^25 triangles, count <= 64)
RL_HD int leaf_ref(uint32_t first, uint32_t count) { return (int)(0x80000000u | ((count - 1u) << 25) | first); }
RL_HD uint32_t leaf_first(int ref) { return (uint32_t)ref & 0x01ffffffu; }
RL_HD uint32_t leaf_count(int ref) { return (((uint32_t)ref >> 25) & 63u) + 1u; }

struct HitRec {
    float t, u, v;
    uint32_t prim;
};
struct Trav {
    V3 o, d;
    V3 inv;        // 1/d (clamped away from 0)
    V3 ood_n, ood_f; // o/d widened for far origins (+- 1e-5 |o| / |d|, 0 otherwise): entry t = fma(near plane, inv, -ood_n), exit t = fma(far plane, inv, -ood_f)
    float rs2, rs8; // 2e-6 and 8e-6 times the coordinate scale of this ray (prefilter margins)
    float tmax;    // closest: best t so far widened by RL_TIE_WINDOW (culling bound); shadow: the segment's threshold
    float best;    // closest: best t so far
    bool amb;      // closest: two accepted hits within the tie window, or the best hit lies on the rim of its triangle -> the reference's
                   // own traversal decides (ref_bvh_closest); shadow: the segment is blocked only by rim hits so far (ref_bvh_any decides)
    bool edge;
    uint32_t slot; // closest: Morton slot of the best triangle
    uint32_t rim_slot; // shadow: Morton slot of the last rim blocker
    float tie_t;   // closest: smallest t of a tie seen so far (-1: none); it matters only when it is within the window of the final hit
    bool edge_checks; // the scene carries the reference's tree (sv.ref_nodes)
    bool wide4;       // sv.wide4
    uint32_t near_off; // 4-wide nodes: which of the float4 rows {lo.x lo.y lo.z hi.x hi.y hi.z} holds the NEAR plane of each axis (2 bits per axis: 0 / 3 + axis);
                       // the far plane is the other one (rl_wide_host.hpp)
    float omax, abs_max; // max |o|, scene extent (hit_unsafe)
    float u, v;
    uint32_t prim;
    int cur;       // >= 0 inner node, < 0 leaf ~cur, RL_TRAV_DONE
    int sp;        // entries in the caller's stack array (kept outside the struct so the state stays in registers)
};
RL_HD float clamp_inv(float inv) {
    // |1/d| > 1e18 (incl. +-inf for d = +-0) would give NaN in the fma form; a finite 1e18 keeps every product finite
    return fabsf(inv) <= 1e18f ? inv : copysignf(1e18f, inv);
}
RL_HD void trav_begin(Trav &tr, const SceneView &sv, V3 o, V3 d, V3 inv, float tmax) {
    tr.o = o;
    tr.d = d;
    tr.inv = V3{clamp_inv(inv.x), clamp_inv(inv.y), clamp_inv(inv.z)};
    const V3 ood = V3{o.x * tr.inv.x, o.y * tr.inv.y, o.z * tr.inv.z};
    float m = fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fabsf(o.z));
    const bool far = m > 8.0f * sv.abs_max;
    tr.omax = m, tr.abs_max = sv.abs_max;
    const float eps = far ? 1e-5f * m : 0.0f; // (an inflation in t far above any rounding of the slab test; exact zero for origins near the scene)
    tr.ood_n = V3{ood.x + eps * fabsf(tr.inv.x), ood.y + eps * fabsf(tr.inv.y), ood.z + eps * fabsf(tr.inv.z)};
    tr.ood_f = V3{ood.x - eps * fabsf(tr.inv.x), ood.y - eps * fabsf(tr.inv.y), ood.z - eps * fabsf(tr.inv.z)};
    tr.tmax = tmax;
    tr.best = tmax;
    tr.amb = false;
    tr.edge = false;
    tr.tie_t = -1.0f;
    tr.edge_checks = RL_REF_ORDER && sv.ref_nodes != nullptr;
    tr.wide4 = sv.wide4 != 0u;
    tr.near_off = (tr.inv.x < 0.0f ? 3u : 0u) | (tr.inv.y < 0.0f ? 3u << 2 : 0u) | (tr.inv.z < 0.0f ? 3u << 4 : 0u);
    tr.u = 0.0f;
    tr.v = 0.0f;
    tr.prim = RL_MISS;
    tr.cur = sv.root_ref;
    tr.sp = 0;
    const float rs = 4.0f * fmaxf(m, sv.abs_max);
    tr.rs2 = 2e-6f * rs;
    tr.rs8 = 8e-6f * rs;
}
// Conservative entry distance of the ray into box (lo,hi) for t in [0, tmax], or -1 on a miss.
// The near / far plane of each axis is picked by the sign of the direction (same values as min / max of the two products: the
// products are monotonic in the plane coordinate), so a box costs 6 fma + 6 selects instead of 6 fma + 6 min/max + 6 adds.
RL_HD float box_entry(const Trav &tr, float lox, float loy, float loz, float hix, float hiy, float hiz) {
    const bool sx = tr.inv.x < 0.0f, sy = tr.inv.y < 0.0f, sz = tr.inv.z < 0.0f;
    const float nx = fmaf(sx ? hix : lox, tr.inv.x, -tr.ood_n.x), fx = fmaf(sx ? lox : hix, tr.inv.x, -tr.ood_f.x);
    const float ny = fmaf(sy ? hiy : loy, tr.inv.y, -tr.ood_n.y), fy = fmaf(sy ? loy : hiy, tr.inv.y, -tr.ood_f.y);
    const float nz = fmaf(sz ? hiz : loz, tr.inv.z, -tr.ood_n.z), fz = fmaf(sz ? loz : hiz, tr.inv.z, -tr.ood_f.z);
    float tmin = fmaxf(fmaxf(nx, ny), fmaxf(nz, 0.0f));
    float tend = fminf(fminf(fx, fy), fminf(fz, tr.tmax));
    return tmin <= tend ? tmin : -1.0f;
}
#ifndef RL_TREE_DIST
#define RL_TREE_DIST 1 // closest-hit rays keep the entry distance of every stacked node and drop it at the pop once a nearer hit is known (A/B hook)
#endif
#ifndef RL_TREE_ANY_NOSORT
#define RL_TREE_ANY_NOSORT 1 // shadow segments do not order the children of a node (A/B hook)
#endif
// Next node from the stack.  sdist (closest-hit rays): the entry distances of the stacked nodes (rounded down) -- an entry behind the
// culling bound tr.tmax, which has shrunk since the push, would only be visited to find all its children culled.
// box_entry with the near / far planes already picked
RL_HD float box_entry_nf(const Trav &tr, float nxp, float nyp, float nzp, float fxp, float fyp, float fzp) {
    const float tn = fmaxf(fmaxf(fmaf(nxp, tr.inv.x, -tr.ood_n.x), fmaf(nyp, tr.inv.y, -tr.ood_n.y)), fmaxf(fmaf(nzp, tr.inv.z, -tr.ood_n.z), 0.0f));
    const float tf = fminf(fminf(fmaf(fxp, tr.inv.x, -tr.ood_f.x), fmaf(fyp, tr.inv.y, -tr.ood_f.y)), fminf(fmaf(fzp, tr.inv.z, -tr.ood_f.z), tr.tmax));
    return tn <= tf ? tn : -1.0f;
}
RL_HD int trav_pop(Trav &tr, const int *stack, const float *sdist = nullptr) {
    if (RL_TREE_DIST && sdist) {
        while (tr.sp > 0) {
            --tr.sp;
            if (sdist[tr.sp] <= tr.tmax) return stack[tr.sp];
        }
        return RL_TRAV_DONE;
    }
    return tr.sp > 0 ? stack[--tr.sp] : RL_TRAV_DONE;
}
// 4-wide node (rl_wide_host.hpp): four slab tests per visit.  The children that are hit are ordered by entry distance with a
// sorting network over keys (distance bits with the child number in the two lowest bits: the order is a heuristic, two mantissa
// bits do not matter); the nearest becomes tr.cur, the others go on the stack farthest first.
// ANY (shadow segments: every node the segment crosses is visited unless a blocker ends the walk, so the order is worth nothing):
// the first child hit becomes tr.cur, the others are pushed as they come.
RL_HD void sort2u(uint32_t &a, uint32_t &b) {
    const uint32_t lo = a < b ? a : b, hi = a < b ? b : a;
    a = lo, b = hi;
}
template <bool ANY>
RL_HD void trav_node_step4(Trav &tr, int *stack, const float4 *nodes, float *sdist) {
    const float4 *nd = nodes + 7 * tr.cur;
    // The near / far plane of each axis is picked by the sign of the direction (box_entry) -- the same for every node of a ray, so the
    // choice sits in the ADDRESS of the row (rows 0-2 lo, 3-5 hi): 6 fma per box and no selects.  Unused slots hold an inverted
    // infinite box and fail the test by themselves.
    const uint32_t ox = tr.near_off & 3u, oy = (tr.near_off >> 2) & 3u, oz = tr.near_off >> 4;
    const float4 nx = nd[ox], ny = nd[1u + oy], nz = nd[2u + oz], fx = nd[3u - ox], fy = nd[4u - oy], fz = nd[5u - oz], rf = nd[6];
    const float d0 = box_entry_nf(tr, nx.x, ny.x, nz.x, fx.x, fy.x, fz.x), d1 = box_entry_nf(tr, nx.y, ny.y, nz.y, fx.y, fy.y, fz.y);
    const float d2 = box_entry_nf(tr, nx.z, ny.z, nz.z, fx.z, fy.z, fz.z), d3 = box_entry_nf(tr, nx.w, ny.w, nz.w, fx.w, fy.w, fz.w);
    const uint32_t c0 = f2u(rf.x), c1 = f2u(rf.y), c2 = f2u(rf.z), c3 = f2u(rf.w);
    if (ANY && RL_TREE_ANY_NOSORT) {
        int next = RL_TRAV_DONE;
        if (d3 >= 0.0f) next = (int)c3;
        if (d2 >= 0.0f) {
            if (next != RL_TRAV_DONE && tr.sp < RL_STACK_SIZE) stack[tr.sp++] = next;
            next = (int)c2;
        }
        if (d1 >= 0.0f) {
            if (next != RL_TRAV_DONE && tr.sp < RL_STACK_SIZE) stack[tr.sp++] = next;
            next = (int)c1;
        }
        if (d0 >= 0.0f) {
            if (next != RL_TRAV_DONE && tr.sp < RL_STACK_SIZE) stack[tr.sp++] = next;
            next = (int)c0;
        }
        tr.cur = next != RL_TRAV_DONE ? next : trav_pop(tr, stack);
        return;
    }
    // d >= 0 on a hit (float bits are monotonic), -1 on a miss -> key 0xffffffff
    uint32_t k0 = d0 >= 0.0f ? (f2u(d0) & ~3u) : 0xffffffffu;
    uint32_t k1 = d1 >= 0.0f ? ((f2u(d1) & ~3u) | 1u) : 0xffffffffu;
    uint32_t k2 = d2 >= 0.0f ? ((f2u(d2) & ~3u) | 2u) : 0xffffffffu;
    uint32_t k3 = d3 >= 0.0f ? ((f2u(d3) & ~3u) | 3u) : 0xffffffffu;
    sort2u(k0, k1), sort2u(k2, k3), sort2u(k0, k2), sort2u(k1, k3), sort2u(k1, k2);
#define RL_PICK4(k) (int)(((k) & 2u) ? (((k) & 1u) ? c3 : c2) : (((k) & 1u) ? c1 : c0))
#define RL_PUSH4(k)                                                             \
    if ((k) != 0xffffffffu && tr.sp < RL_STACK_SIZE) {                          \
        if (RL_TREE_DIST && sdist) sdist[tr.sp] = u2f((k) & ~3u); /* <= d */    \
        stack[tr.sp++] = RL_PICK4(k);                                           \
    }
    if (k0 == 0xffffffffu) {
        tr.cur = trav_pop(tr, stack, sdist);
        return;
    }
    RL_PUSH4(k3)
    RL_PUSH4(k2)
    RL_PUSH4(k1)
    tr.cur = RL_PICK4(k0);
#undef RL_PUSH4
#undef RL_PICK4
}
// Phase A: one inner-node step (tr.cur >= 0).  Leaves tr.cur at a child, a popped entry or DONE.
template <bool ANY = false>
RL_HD void trav_node_step(Trav &tr, int *stack, const float4 *nodes, float *sdist = nullptr) {
    if (tr.wide4) {
        trav_node_step4<ANY>(tr, stack, nodes, sdist);
        return;
    }
    const int node = tr.cur;
    float4 a = nodes[4 * node + 0], b = nodes[4 * node + 1], c = nodes[4 * node + 2], k = nodes[4 * node + 3];
    float d0 = box_entry(tr, a.x, a.y, a.z, a.w, b.x, b.y);
    float d1 = box_entry(tr, b.z, b.w, c.x, c.y, c.z, c.w);
    int c0 = (int)f2u(k.x), c1 = (int)f2u(k.y);
    if (d0 >= 0.0f && d1 >= 0.0f) {
        bool swap = d1 < d0;
        int nearc = swap ? c1 : c0, farc = swap ? c0 : c1;
        if (tr.sp < RL_STACK_SIZE) {
            if (RL_TREE_DIST && sdist) sdist[tr.sp] = swap ? d0 : d1;
            stack[tr.sp++] = farc;
        }
        tr.cur = nearc;
    } else if (d0 >= 0.0f) tr.cur = c0;
    else if (d1 >= 0.0f) tr.cur = c1;
    else tr.cur = trav_pop(tr, stack, sdist);
}
// Conservative prefilter in front of the exact triangle test: approximate hit parameter and
// barycentrics from the precomputed plane / affine functionals (fma, reciprocal), with error
// margins that dominate every rounding term of both this estimate and the reference's own
// arithmetic.  Returns false only when the exact test is certain to reject.
// NaNs (degenerate triangles, d.n == 0) fall through to the exact test.
RL_HD float rcp_fast(float x) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); // one MUFU.RCP; 1 ulp, covered by the margins
    return r;
#else
    return 1.0f / x;
#endif
}
RL_HD int ffs64(uint64_t m) {
#if defined(__CUDA_ARCH__)
    return __ffsll((long long)m);
#else
    return __builtin_ffsll((long long)m);
#endif
}
RL_HD bool tri_prefilter(const Trav &tr, float4 r2, float4 r3, float4 r4, float4 r5) {
    // plane: t' = (pn - o.n) / (d.n); |t' - t| <= mt covers the roundings of both this estimate and
    // the exact path (|d.n| <= 1, so |1/(d.n)| >= 1 and mt >= 1e-6 |t'|)
    float den = fmaf(tr.d.x, r3.x, fmaf(tr.d.y, r3.y, tr.d.z * r3.z));
    float on = fmaf(tr.o.x, r3.x, fmaf(tr.o.y, r3.y, tr.o.z * r3.z));
    float num = r3.w - on;
    float rden = rcp_fast(den);
    float tp = num * rden;
    float ard = fabsf(rden);
    float mt = ard * fmaf(fabsf(num), fmaf(ard, 1e-6f, 4e-6f), tr.rs2);
    // barycentrics at p' = o + t' d from the affine functionals; margin m >= |Mu| (4 mt + 8e-6 rs + 2e-6 |t'|)
    float px = fmaf(tp, tr.d.x, tr.o.x), py = fmaf(tp, tr.d.y, tr.o.y), pz = fmaf(tp, tr.d.z, tr.o.z);
    float up = fmaf(px, r4.x, fmaf(py, r4.y, fmaf(pz, r4.z, r4.w)));
    float vp = fmaf(px, r5.x, fmaf(py, r5.y, fmaf(pz, r5.z, r5.w)));
    float m = fmaf(r2.w, fmaf(mt, 6.0f, tr.rs8), 1e-6f);
    // branch-free: every comparison is false for NaN, so degenerate cases fall through to the exact test
    bool reject = (tp < -mt) | (tp > tr.tmax + mt) | (fminf(up, vp) < -m) | (up + vp > 1.0f + m);
    return !reject;
}
// Phase B (closest hit): tr.cur is a leaf reference.  Two uniform loops: the scan runs only the
// prefilter and records the survivors in a bit mask (a leaf holds at most 64 triangles); the
// exact tests then run back to back, so that the long exact path is not executed once per scan
// iteration for the one or two lanes that need it.
RL_HD uint64_t leaf_scan(const Trav &tr, const float4 *trav, uint32_t first, uint32_t count) {
    uint64_t mask = 0;
    for (uint32_t k = 0; k < count; k++) {
        const float4 *r = trav + 6 * (first + k);
        if (tri_prefilter(tr, r[2], r[3], r[4], r[5])) mask |= 1ull << k;
    }
    return mask;
}
RL_HD void trav_leaf_closest(Trav &tr, const int *stack, const float4 *trav, const float *sdist = nullptr) {
    const uint32_t first = leaf_first(tr.cur), count = leaf_count(tr.cur);
    uint64_t mask = leaf_scan(tr, trav, first, count);
    while (mask) {
        const uint32_t k = (uint32_t)ffs64(mask) - 1u;
        mask &= mask - 1;
        const float4 *r = trav + 6 * (first + k);
        float4 r1 = r[1];
        float t_, u_, v_;
        if (tri_test(r[0], r1, r[2], r[3], tr.o, tr.d, tr.tmax, &t_, &u_, &v_)) {
            const uint32_t pw = f2u(r1.w), prim_ = pw & RL_PRIM_MASK;
            if (tr.prim != RL_MISS && prim_ != tr.prim && tie_window(t_, tr.best)) tr.tie_t = tr.tie_t < 0.0f ? fminf(t_, tr.best) : fminf(tr.tie_t, fminf(t_, tr.best));
            // without the reference's tree (sv.ref_nodes == nullptr): strict `t < its.t` in mesh-major order => on exact ties the lowest index wins
            if (t_ < tr.best || (t_ == tr.best && tr.prim != RL_MISS && prim_ < tr.prim)) {
                tr.best = t_;
                tr.slot = first + k;
                tr.edge = hit_near_edge(u_, v_, r[2].w, tr.rs8) || hit_unsafe(pw, t_, tr.omax, tr.abs_max);
                tr.tmax = tie_bound(t_);
                tr.u = u_;
                tr.v = v_;
                tr.prim = prim_;
            }
        }
    }
    tr.cur = trav_pop(tr, stack, sdist);
}
// Phase B (any hit): returns true when the segment is blocked.
RL_HD bool trav_leaf_any(Trav &tr, const int *stack, const float4 *trav) {
    const uint32_t first = leaf_first(tr.cur), count = leaf_count(tr.cur);
    uint64_t mask = leaf_scan(tr, trav, first, count);
    while (mask) {
        const uint32_t k = (uint32_t)ffs64(mask) - 1u;
        mask &= mask - 1;
        const float4 *r = trav + 6 * (first + k);
        float t_, u_, v_;
        const float4 r1 = r[1];
        if (tri_test(r[0], r1, r[2], r[3], tr.o, tr.d, tr.tmax, &t_, &u_, &v_) && t_ < tr.tmax) {
            if (tr.edge_checks && (hit_near_edge(u_, v_, r[2].w, tr.rs8) || hit_unsafe(f2u(r1.w), t_, tr.omax, tr.abs_max))) {
                tr.rim_slot = first + k; // a rim hit: the caller checks whether the reference reaches it (ref_path_ok) ...
                tr.amb = true;           // ... meanwhile keep looking for a blocker the reference cannot miss
                continue;
            }
            tr.cur = RL_TRAV_DONE;
            return true;
        }
    }
    tr.cur = trav_pop(tr, stack);
    return false;
}

// BVHAccel::intersect (accel.rs:243-288) over the reference's own tree, for a ray that passed the root test (tnear = 1e-4; `tfar`
// and the initial its.t are f32::MAX for trace(), the shadow threshold for visible()).  The recursion becomes a stack of
// (node, entry distance): the far child waits on the stack and its `d2 < its.t` test runs when it is popped, i.e. after the near
// subtree has finished -- exactly where the reference evaluates it.  Leaves test their (<= 2) primitives in order with the
// reference's strict `t < its.t`.  ANY: stop at the first accepted triangle (visible() only asks whether there is one).  The
// tree's depth is checked against RL_STACK_SIZE when the scene is built.  Returns true when a triangle was accepted.
template <bool ANY>
RL_HD bool ref_bvh_walk(const float4 *ref_nodes, const uint32_t *ref_prims, const float4 *trav, V3 o, V3 d, float tfar, float *t_io, float *u_io, float *v_io, uint32_t *prim_io) {
    const V3 inv = V3{1.0f / d.x, 1.0f / d.y, 1.0f / d.z};
    float best = tfar, bu = 0.0f, bv = 0.0f;
    uint32_t bprim = RL_MISS;
    int st_node[RL_STACK_SIZE];
    float st_dist[RL_STACK_SIZE];
    int sp = 0;
    int node = 0;
    for (;;) {
        const float4 n0 = ref_nodes[2 * node], n1 = ref_nodes[2 * node + 1];
        const uint32_t info = f2u(n0.w), count = f2u(n1.w);
        if (count != 0u) {
            for (uint32_t k = 0; k < count; k++) {
                const float4 *r = trav + 6 * ref_prims[info + k];
                const float4 r1 = r[1];
                float t_, u_, v_;
                if (tri_test(r[0], r1, r[2], r[3], o, d, best, &t_, &u_, &v_) && t_ < best) {
                    best = t_, bu = u_, bv = v_, bprim = f2u(r1.w) & RL_PRIM_MASK;
                    if (ANY) return true;
                }
            }
        } else {
            int id1 = (int)info, id2 = (int)info + 1;
            float d1, d2;
            const float4 a0 = ref_nodes[2 * id1], a1 = ref_nodes[2 * id1 + 1], b0 = ref_nodes[2 * id2], b1 = ref_nodes[2 * id2 + 1];
            if (!aabb_entry_ref(xyz(a0), xyz(a1), o, inv, RL_EPSILON, tfar, &d1)) d1 = u2f(0x7f800000u);
            if (!aabb_entry_ref(xyz(b0), xyz(b1), o, inv, RL_EPSILON, tfar, &d2)) d2 = u2f(0x7f800000u);
            if (d1 > d2) {
                const float td = d1;
                d1 = d2, d2 = td;
                const int ti = id1;
                id1 = id2, id2 = ti;
            }
            if (sp < RL_STACK_SIZE) st_node[sp] = id2, st_dist[sp] = d2, sp++;
            if (d1 < best) {
                node = id1;
                continue;
            }
        }
        bool found = false;
        while (sp > 0) {
            sp--;
            if (st_dist[sp] < best) {
                node = st_node[sp];
                found = true;
                break;
            }
        }
        if (!found) break;
    }
    if (!ANY) *t_io = bprim == RL_MISS ? RL_F32_MAX : best, *u_io = bu, *v_io = bv, *prim_io = bprim;
    return bprim != RL_MISS;
}
RL_HD_NOINLINE void ref_bvh_closest_impl(const float4 *ref_nodes, const uint32_t *ref_prims, const float4 *trav, V3 o, V3 d, float *t_io, float *u_io, float *v_io,
                                         uint32_t *prim_io) {
    ref_bvh_walk<false>(ref_nodes, ref_prims, trav, o, d, RL_F32_MAX, t_io, u_io, v_io, prim_io);
}
RL_HD_NOINLINE bool ref_bvh_any_impl(const float4 *ref_nodes, const uint32_t *ref_prims, const float4 *trav, V3 o, V3 d, float thr) {
    float t_, u_, v_;
    uint32_t p_;
    return ref_bvh_walk<true>(ref_nodes, ref_prims, trav, o, d, thr, &t_, &u_, &v_, &p_);
}
RL_HD void ref_bvh_closest(const SceneView &sv, const float4 *trav, V3 o, V3 d, float *t_io, float *u_io, float *v_io, uint32_t *prim_io) {
#ifdef RL_REF_NOCALL
    *t_io = *t_io + 0.0f * (float)(uintptr_t)sv.ref_nodes;
    return;
#endif
    ref_bvh_closest_impl(sv.ref_nodes, sv.ref_prims, trav, o, d, t_io, u_io, v_io, prim_io);
}
// Does the reference reach the leaf of the triangle at Morton slot `slot`?  It does when every box on the way up to the root
// passes its slab test (the caller has tested the root): the hit lies in all of them, so their entry distances are <= t (1 + ulps),
// below any its.t the reference can hold before it finds this hit (the hit is the nearest by more than the tie window).  This settles
// rim hits and hit_unsafe hits without a full walk; what fails here goes to ref_bvh_closest / ref_bvh_any.
RL_HD bool ref_path_ok(const SceneView &sv, uint32_t slot, V3 o, V3 d, float tfar) {
    const V3 inv = V3{1.0f / d.x, 1.0f / d.y, 1.0f / d.z};
    uint32_t node = sv.ref_up[sv.ref_n_nodes + slot];
    while (node != 0u) {
        const float4 n0 = sv.ref_nodes[2 * node], n1 = sv.ref_nodes[2 * node + 1];
        float t_entry;
        if (!aabb_entry_ref(xyz(n0), xyz(n1), o, inv, RL_EPSILON, tfar, &t_entry)) return false;
        node = sv.ref_up[node];
    }
    return true;
}
RL_HD bool ref_bvh_any(const SceneView &sv, const float4 *trav, V3 o, V3 d, float thr) {
#ifdef RL_REF_NOCALL
    return thr > 1e30f;
#endif
    return ref_bvh_any_impl(sv.ref_nodes, sv.ref_prims, trav, o, d, thr);
}
// ---- flat quad scan (scenes of a few dozen triangles) -----------------------------------------------
// Incoherent rays gain nothing from a hierarchy over ~36 triangles (every lane of a warp walks a different
// branch), so they run a conservative prefilter over ALL triangles in lockstep and the exact test only on
// the survivors.  To make the scan cheap the triangles are grouped at build time (rl_flat_host.hpp):
//   quad record  = two triangles A, B lying in one plane (a quad face): ONE ray/plane intersection with A's plane
//                  (B's vertices are within flat_delta of it), ONE point-in-parallelogram test with two affine
//                  functionals U, V in [0, 1] over A u B, and -- when A and B lie on opposite sides of a shared
//                  edge -- the functional D = k0 + kU U + kV V (>= 0 on A, <= 0 on B) that tells which of the two
//                  the ray can hit; unpaired triangles get a record of their own (B's bit masked by flat_valid_b);
//   group        = two quad records P, Q interleaved component-wise, so that every arithmetic step is one
//                  packed fma.rn.f32x2 / add / mul (FFMA2 / FADD2 / FMUL2 on sm_100) over (P, Q).
// Eight float4 per group (two quads = up to four triangles):
//   [0] {n.x P,Q  n.y P,Q}   [1] {n.z P,Q  pn P,Q}            plane of A: n.p = pn
//   [2] {MU.x P,Q MU.y P,Q}  [3] {MU.z P,Q cU P,Q}             U(p) = MU.p + cU           [4],[5] the same for V
//   [6] {mn P,Q  k0 P,Q}     [7] {kU P,Q  kV P,Q}              mn = largest gradient norm of U, V, D
// followed (after the last group) by 64 bytes: Morton slot (index into trav[]) of A [bit] and of B [32 + bit].
// The scan yields three sign bits per quad, shifted into three masks in scan order (first quad ends in the highest
// bit): q = outside the parallelogram or outside the t range, dp = D + m < 0 (not A), dm = m - D < 0 (not B).
// Rejection is decided by the sign of a minimum: FMNMX drops NaN operands and an all-NaN minimum is the canonical
// (positive) NaN, so degenerate cases (d.n == 0, inf arithmetic, NaN functionals) stay candidates.
#define RL_FLAT_F4 8
#define RL_FLAT_TAIL_F4 4
#define RL_FLAT_MAX_GROUPS 16
struct F2 {
    float x, y;
};
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ unsigned long long f2_pack(F2 a) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y));
    return r;
}
__device__ __forceinline__ F2 f2_unpack(unsigned long long v) {
    F2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
#endif
RL_HD F2 f2(float x, float y) { return F2{x, y}; }
RL_HD F2 f2b(float x) { return F2{x, x}; }
RL_HD F2 fma2(F2 a, F2 b, F2 c) {
#if defined(__CUDA_ARCH__)
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(f2_pack(a)), "l"(f2_pack(b)), "l"(f2_pack(c)));
    return f2_unpack(r);
#else
    return F2{fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)};
#endif
}
RL_HD F2 mul2(F2 a, F2 b) {
#if defined(__CUDA_ARCH__)
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_pack(a)), "l"(f2_pack(b)));
    return f2_unpack(r);
#else
    return F2{a.x * b.x, a.y * b.y};
#endif
}
RL_HD F2 add2(F2 a, F2 b) {
#if defined(__CUDA_ARCH__)
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_pack(a)), "l"(f2_pack(b)));
    return f2_unpack(r);
#else
    return F2{a.x + b.x, a.y + b.y};
#endif
}
RL_HD F2 sub2(F2 a, F2 b) {
#if defined(__CUDA_ARCH__)
    unsigned long long r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_pack(a)), "l"(f2_pack(b)));
    return f2_unpack(r);
#else
    return F2{a.x - b.x, a.y - b.y};
#endif
}
// min of three that ignores NaN operands (FMNMX3 on sm_100)
RL_HD float min3f(float a, float b, float c) { return fminf(fminf(a, b), c); }
// (mask << 1) | reject, reject = "s is negative and not NaN"
RL_HD uint32_t push_reject(uint32_t mask, float s) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(__float_as_uint(s), mask, 1); // NaN results are the canonical 0x7fffffff: sign clear
#else
    return (mask << 1) | ((s < 0.0f) ? 1u : 0u);
#endif
}
struct FlatRay {
    V3 o, d;
    float tmax;
    float rs2, rs8; // margins scaled by the coordinate magnitude of this ray (as in trav_begin) + flat_delta
    float omax;     // max |o|
};
RL_HD FlatRay flat_ray(const SceneView &sv, V3 o, V3 d, float tmax) {
    FlatRay fr;
    fr.o = o, fr.d = d, fr.tmax = tmax;
    float m = fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fabsf(o.z));
    const float rs = 4.0f * fmaxf(m, sv.abs_max);
    fr.omax = m;
    fr.rs2 = RL_FLAT_MARGIN_SCALE * (3e-6f * rs + 2.0f * sv.flat_delta);
    fr.rs8 = RL_FLAT_MARGIN_SCALE * 8e-6f * rs;
    return fr;
}
struct FlatMasks {
    uint32_t q, dp, dm;
};
// The six sign values of one group (two quads P, Q): outside parallelogram / t range, not A, not B.
struct FlatSigns {
    float qP, qQ, dpP, dpQ, dmP, dmQ;
};
// TMAX = false: closest-hit rays (tmax = f32::MAX: the far end of the t range cannot reject).
template <bool TMAX>
RL_HD FlatSigns flat_group(const FlatRay &fr, const float4 *g, F2 dx, F2 dy, F2 dz, F2 ox, F2 oy, F2 oz, F2 nox, F2 noy, F2 noz) {
    const float4 a0 = g[0], a1 = g[1];
    const F2 nx = f2(a0.x, a0.y), ny = f2(a0.z, a0.w), nz = f2(a1.x, a1.y), pn = f2(a1.z, a1.w);
    // plane: t' = (pn - o.n) / (d.n)
    const F2 den = fma2(dx, nx, fma2(dy, ny, mul2(dz, nz)));
    const F2 num = fma2(nox, nx, fma2(noy, ny, fma2(noz, nz, pn)));
    const F2 rden = f2(rcp_fast(den.x), rcp_fast(den.y));
    const F2 tp = mul2(num, rden);
    const float ard0 = fabsf(rden.x), ard1 = fabsf(rden.y);
    const F2 mt = f2(ard0 * fmaf(fabsf(num.x), fmaf(ard0, RL_FLAT_MARGIN_SCALE * 2e-6f, RL_FLAT_MARGIN_SCALE * 8e-6f), fr.rs2), ard1 * fmaf(fabsf(num.y), fmaf(ard1, RL_FLAT_MARGIN_SCALE * 2e-6f, RL_FLAT_MARGIN_SCALE * 8e-6f), fr.rs2));
    const F2 px = fma2(tp, dx, ox), py = fma2(tp, dy, oy), pz = fma2(tp, dz, oz);
    const float4 c6 = g[6], c7 = g[7];
    // margin of the in-plane functionals: m >= |grad| (6 mt + 8e-6 rs) + 2e-6
    const F2 m = fma2(f2(c6.x, c6.y), fma2(mt, f2b(6.0f), f2b(fr.rs8)), f2b(RL_FLAT_MARGIN_SCALE * 2e-6f));
    const F2 onem = add2(f2b(1.0f), m);
    const F2 q1 = add2(tp, mt);
    const float4 b0 = g[2], b1 = g[3], b2 = g[4], b3 = g[5];
    const F2 U = fma2(px, f2(b0.x, b0.y), fma2(py, f2(b0.z, b0.w), fma2(pz, f2(b1.x, b1.y), f2(b1.z, b1.w))));
    const F2 V = fma2(px, f2(b2.x, b2.y), fma2(py, f2(b2.z, b2.w), fma2(pz, f2(b3.x, b3.y), f2(b3.z, b3.w))));
    const F2 ua = add2(U, m), va = add2(V, m), ub = sub2(onem, U), vb = sub2(onem, V);
    const F2 D = fma2(f2(c7.x, c7.y), U, fma2(f2(c7.z, c7.w), V, f2(c6.z, c6.w)));
    const F2 Dp = add2(D, m), Dm = sub2(m, D);
    FlatSigns sg;
    if (TMAX) {
        const F2 q2 = sub2(add2(f2b(fr.tmax), mt), tp);
        sg.qP = fminf(min3f(ua.x, va.x, ub.x), min3f(vb.x, q1.x, q2.x));
        sg.qQ = fminf(min3f(ua.y, va.y, ub.y), min3f(vb.y, q1.y, q2.y));
    } else {
        sg.qP = min3f(min3f(ua.x, va.x, ub.x), vb.x, q1.x);
        sg.qQ = min3f(min3f(ua.y, va.y, ub.y), vb.y, q1.y);
    }
    sg.dpP = Dp.x, sg.dpQ = Dp.y, sg.dmP = Dm.x, sg.dmQ = Dm.y;
    return sg;
}
RL_HD uint32_t sign_bit(float s) { return f2u(s) >> 31; } // canonical NaNs are positive
// Scan the groups.  CULL (camera rays): `quads` = the quads that overlap the frustum of the warp's pixels (rl_kernels.cuh:
// k_camera_cull), uniform over the warp: only groups with a quad in it are scanned, everything else counts as rejected.
#ifndef RL_SCAN_UNROLL
#define RL_SCAN_UNROLL 1
#endif
constexpr int kScanUnroll = RL_SCAN_UNROLL;
template <bool CULL, bool TMAX>
RL_HD FlatMasks flat_scan(const FlatRay &fr, const float4 *flat, uint32_t n_groups, uint32_t quads) {
    FlatMasks mk;
    const F2 dx = f2b(fr.d.x), dy = f2b(fr.d.y), dz = f2b(fr.d.z);
    const F2 ox = f2b(fr.o.x), oy = f2b(fr.o.y), oz = f2b(fr.o.z);
    const F2 nox = f2b(-fr.o.x), noy = f2b(-fr.o.y), noz = f2b(-fr.o.z);
    if (CULL) {
        mk.q = 0xffffffffu, mk.dp = 0u, mk.dm = 0u;
        uint32_t todo = (quads | (quads >> 1)) & 0x55555555u; // bit 2k: group with quad bits 2k, 2k+1 has work
        while (todo) {
            const uint32_t lo = (uint32_t)ffs64((uint64_t)todo) - 1u; // bit position of the group's Q quad (P = lo + 1)
            todo &= todo - 1u;
            const uint32_t gi = n_groups - 1u - (lo >> 1);
            const FlatSigns sg = flat_group<TMAX>(fr, flat + RL_FLAT_F4 * gi, dx, dy, dz, ox, oy, oz, nox, noy, noz);
            mk.q = (mk.q & ~(3u << lo)) | (((sign_bit(sg.qP) << 1) | sign_bit(sg.qQ)) << lo);
            mk.dp |= ((sign_bit(sg.dpP) << 1) | sign_bit(sg.dpQ)) << lo;
            mk.dm |= ((sign_bit(sg.dmP) << 1) | sign_bit(sg.dmQ)) << lo;
        }
        return mk;
    }
    mk.q = 0u, mk.dp = 0u, mk.dm = 0u;
#pragma unroll kScanUnroll
    for (uint32_t gi = 0; gi < n_groups; gi++) {
        const FlatSigns sg = flat_group<TMAX>(fr, flat + RL_FLAT_F4 * gi, dx, dy, dz, ox, oy, oz, nox, noy, noz);
        mk.q = push_reject(push_reject(mk.q, sg.qP), sg.qQ);
        mk.dp = push_reject(push_reject(mk.dp, sg.dpP), sg.dpQ);
        mk.dm = push_reject(push_reject(mk.dm, sg.dmP), sg.dmQ);
    }
    return mk;
}
// Candidate triangles: bits [0, 32) the A triangles, bits [32, 64) the B triangles of the quads (bit b <-> scan index 2 n_groups - 1 - b).
template <bool CULL, bool TMAX>
RL_HD uint64_t flat_candidates(const FlatRay &fr, const SceneView &sv, const float4 *flat, uint32_t quads) {
    const FlatMasks mk = flat_scan<CULL, TMAX>(fr, flat, sv.n_groups, quads);
    uint32_t ca = ~(mk.q | mk.dp) & sv.flat_valid_a, cb = ~(mk.q | mk.dm) & sv.flat_valid_b;
    if (CULL) ca &= quads, cb &= quads;
    return ((uint64_t)cb << 32) | (uint64_t)ca;
}
RL_HD uint32_t flat_slot(const SceneView &sv, const float4 *flat, uint32_t bit) {
    return (uint32_t)reinterpret_cast<const unsigned char *>(flat + RL_FLAT_F4 * sv.n_groups)[bit];
}
// Mesh::intersection_tri with the square roots and divisions of (u, v) deferred.  All decisions of tri_test are taken,
// in the same arithmetic, up to the last one (u + v <= 1, which also implies u, v <= 1): that one is decided from
// alpha = (e1 x pv).n ~ det v and beta = (pv x e2).n ~ det u when alpha + beta is clear of det by the margin `ml`
// (|u0| <= |alpha| / |n| + |e1| eta with eta the out-of-plane rounding of p, far below ml det: DESIGN.md section 6),
// and by the reference's own (u, v) otherwise.  Returns 0 = rejected, 1 = accepted with (u, v) still to be computed
// by tri_uv, 2 = accepted and (u, v) computed.
RL_HD int tri_test_lazy(float4 r0, float4 r1, float4 r2, float4 r3, V3 o, V3 d, float t_bound, float rs8, float *t_out, float *u_out, float *v_out, bool *near_edge) {
    V3 v0 = xyz(r0), e1 = xyz(r1), e2 = xyz(r2), n_geo = xyz(r3);
    float det = r0.w;
    float denom = dot(d, n_geo);
    if (denom == 0.0f) return 0;
    float t = -dot(o - v0, n_geo) / denom;
    if (t < 0.0f) return 0;
    if (!(t <= t_bound) || !(t > 0.00001f)) return 0;
    V3 p = o + t * d;
    V3 pv = p - v0;
    V3 u0 = cross(e1, pv);
    float alpha = dot(u0, n_geo);
    if (alpha < 0.0f) return 0;
    V3 v0c = cross(pv, e2);
    float beta = dot(v0c, n_geo);
    if (beta < 0.0f) return 0;
    *t_out = t;
    const float ml = fmaf(r2.w, rs8, RL_FLAT_MARGIN_SCALE * 2e-6f);
    const float s = alpha + beta;
    if (near_edge) { // not strictly interior by the rim margin (see hit_near_edge)
        const float mr = rim_margin(r2.w, rs8);
        *near_edge = !(fminf(alpha, beta) >= det * mr && s <= det * (1.0f - mr));
    }
    if (s <= det * (1.0f - ml)) return 1; // false for NaN
    if (s > det * (1.0f + ml)) return 0;
    float v = magnitude(u0) / det;
    float u = magnitude(v0c) / det;
    if (u < 0.0f || v < 0.0f || u > 1.0f || v > 1.0f) return 0;
    if (!(u + v <= 1.0f)) return 0;
    *u_out = u;
    *v_out = v;
    return 2;
}
// (u, v) of an accepted hit: the operations of tri_test on the same inputs, hence the same bits
RL_HD void tri_uv(float4 r0, float4 r1, float4 r2, V3 o, V3 d, float t, float *u_out, float *v_out) {
    V3 v0 = xyz(r0), e1 = xyz(r1), e2 = xyz(r2);
    float det = r0.w;
    V3 p = o + t * d;
    V3 pv = p - v0;
    V3 u0 = cross(e1, pv);
    V3 v0c = cross(pv, e2);
    *v_out = magnitude(u0) / det;
    *u_out = magnitude(v0c) / det;
}
// Closest hit over the flat table: same result rule as trav_leaf_closest.  The loop only keeps the best and the second-best accepted
// t; whether the winner is one the reference finds for certain (not tied, not on the rim, not hit_unsafe) is decided once, after it.
// DEFER: do not walk the reference's tree here; report the ray in *needs_ref instead (the wavefront kernels collect such rays in a list
// that k_fix_flat re-traces right after: a call to the slow path inside the hot kernels costs them 10-20 % through register allocation
// alone, measured).
template <bool CULL = false, bool DEFER = false>
RL_HD HitRec flat_closest(const SceneView &sv, const float4 *flat, const float4 *trav, V3 o, V3 d, uint32_t quads = 0xffffffffu, bool *needs_ref = nullptr) {
    FlatRay fr = flat_ray(sv, o, d, RL_F32_MAX);
    HitRec h;
    h.t = RL_F32_MAX, h.u = 0.0f, h.v = 0.0f, h.prim = RL_MISS;
    uint64_t c = flat_candidates<CULL, false>(fr, sv, flat, quads);
    uint32_t best_slot = 0u;
    bool need_uv = false;
    float t2 = u2f(0x7f800000u); // second smallest accepted t (a different triangle than the best)
    while (c) {
        const uint32_t b = (uint32_t)ffs64(c) - 1u;
        c &= c - 1;
        const uint32_t slot = flat_slot(sv, flat, b);
        const float4 *r = trav + 6 * slot;
        float4 r1 = r[1];
        float t_, u_, v_;
        const int k = tri_test_lazy(r[0], r1, r[2], r[3], o, d, tie_bound(h.t), fr.rs8, &t_, &u_, &v_, nullptr);
        if (k) {
            const uint32_t prim_ = f2u(r1.w) & RL_PRIM_MASK;
            if (t_ < h.t || (t_ == h.t && h.prim != RL_MISS && prim_ < h.prim)) {
                if (h.prim != RL_MISS) t2 = h.t;
                h.t = t_, h.prim = prim_, best_slot = slot, need_uv = k == 1;
                if (k == 2) h.u = u_, h.v = v_;
            } else t2 = fminf(t2, t_);
        }
    }
    if (h.prim == RL_MISS) return h;
    const float4 *rb = trav + 6 * best_slot;
    const float4 rb1 = rb[1], rb2 = rb[2];
    if (need_uv) tri_uv(rb[0], rb1, rb2, o, d, h.t, &h.u, &h.v);
    if (RL_REF_ORDER && sv.ref_nodes) {
        // (nearly) tied hits, a hit on the rim of its triangle or one the reference's boxes may lose: the reference's own traversal decides
        // ties, and hits nearer than the reference's tnear = 1e-4 (every box entry distance is clamped to tnear there, so `d < its.t` culls
        // whole subtrees as soon as ANY such hit is found): the full walk.  Rim / hit_unsafe hits: only when the reference misses the leaf.
        bool walk = t2 <= tie_bound(h.t) || !(h.t > 1.001e-4f);
        if (!walk && (hit_near_edge(h.u, h.v, rb2.w, fr.rs8) || hit_unsafe(f2u(rb1.w), h.t, fr.omax, sv.abs_max))) walk = !ref_path_ok(sv, best_slot, o, d, RL_F32_MAX);
        if (walk) {
            if (DEFER) *needs_ref = true;
            else ref_bvh_closest(sv, trav, o, d, &h.t, &h.u, &h.v, &h.prim);
        }
    }
    return h;
}
// Any hit with t < thr over the flat table (Acceleration::visible): true when the segment is blocked.
template <bool DEFER = false>
RL_HD bool flat_any(const SceneView &sv, const float4 *flat, const float4 *trav, V3 o, V3 d, float thr, bool *needs_ref = nullptr) {
    FlatRay fr = flat_ray(sv, o, d, thr);
    uint64_t c = flat_candidates<false, true>(fr, sv, flat, 0xffffffffu);
    bool rim = false; // blocked only by hits on the rim of their triangle: the reference's boxes may cull them
    while (c) {
        const uint32_t b = (uint32_t)ffs64(c) - 1u;
        c &= c - 1;
        const uint32_t slot = flat_slot(sv, flat, b);
        const float4 *r = trav + 6 * slot;
        float t_, u_, v_;
        bool ne;
        const float4 r1 = r[1];
        if (tri_test_lazy(r[0], r1, r[2], r[3], o, d, thr, fr.rs8, &t_, &u_, &v_, &ne) && t_ < thr) {
            if (!RL_REF_ORDER || !sv.ref_nodes || !(ne || hit_unsafe(f2u(r1.w), t_, fr.omax, sv.abs_max))) return true;
            if (ref_path_ok(sv, slot, o, d, thr)) return true; // the reference reaches this blocker
            rim = true;
        }
    }
    if (rim) {
        if (DEFER) {
            *needs_ref = true; // undecided: k_fix_flat walks the reference's tree and adds the contribution when the segment is visible
            return true;
        }
        return ref_bvh_any(sv, trav, o, d, thr);
    }
    return false;
}

// Acceleration::trace (accel.rs:292-315) without fill_intersection.  Returns false when the
// reference's root-box test rejects the ray (no traversal needed).
RL_HD bool closest_begin(Trav &tr, const SceneView &sv, V3 o, V3 d) {
    V3 inv = V3{1.0f / d.x, 1.0f / d.y, 1.0f / d.z};
    trav_begin(tr, sv, o, d, inv, RL_F32_MAX);
    if (!aabb_intersect_ref(sv.root_min, sv.root_max, o, inv, RL_EPSILON, RL_F32_MAX)) {
        tr.cur = RL_TRAV_DONE;
        return false;
    }
    return true;
}
// Must the reference's own walk decide this ray?  Ties: yes.  Rim / hit_unsafe hits: only when the reference does not reach the leaf.
RL_HD bool closest_ambiguous(const Trav &tr, const SceneView &sv) {
    if (tr.prim == RL_MISS) return false;
    if ((tr.tie_t >= 0.0f && tie_window(tr.tie_t, tr.best)) || !(tr.best > 1.001e-4f)) return true; // (see flat_closest)
    return tr.edge && !ref_path_ok(sv, tr.slot, tr.o, tr.d, RL_F32_MAX);
}
RL_HD HitRec closest_result(const Trav &tr) {
    HitRec h;
    h.t = tr.prim == RL_MISS ? RL_F32_MAX : tr.best;
    h.u = tr.u;
    h.v = tr.v;
    h.prim = tr.prim;
    return h;
}
// Acceleration::visible (accel.rs:316-343): segment setup.  *decided is set when the root test
// already answers (then *vis holds the answer).
// Segment direction, threshold and the reference's root test; false = "not visible" without traversal.
RL_HD bool visible_setup(const SceneView &sv, V3 p0, V3 p1, V3 *d_out, float *thr_out) {
    const float SHADOW_EPS = 0.00001f;
    V3 d = p1 - p0;
    float length = magnitude(d);
    d = d / length;
    float thr = length * (1.0f - SHADOW_EPS);
    V3 inv = V3{1.0f / d.x, 1.0f / d.y, 1.0f / d.z};
    *d_out = d;
    *thr_out = thr;
    return aabb_intersect_ref(sv.root_min, sv.root_max, p0, inv, RL_EPSILON, thr);
}
RL_HD void visible_begin(Trav &tr, const SceneView &sv, V3 p0, V3 p1, bool *decided, bool *vis) {
    const float SHADOW_EPS = 0.00001f;
    V3 d = p1 - p0;
    float length = magnitude(d);
    d = d / length;
    float thr = length * (1.0f - SHADOW_EPS);
    V3 inv = V3{1.0f / d.x, 1.0f / d.y, 1.0f / d.z};
    trav_begin(tr, sv, p0, d, inv, thr);
    *decided = false;
    *vis = true;
    if (!aabb_intersect_ref(sv.root_min, sv.root_max, p0, inv, RL_EPSILON, thr)) {
        tr.cur = RL_TRAV_DONE;
        *decided = true;
        *vis = false; // accel.rs:338-340
    }
}

// Serial drivers (used by the CPU emulator and the small batch kernels; the wavefront kernels
// run the same steps inside a persistent while-while loop, rl_kernels.cuh).
RL_HD HitRec trace_closest(const SceneView &sv, const float4 *flat, const float4 *nodes, const float4 *trav, V3 o, V3 d) {
    Trav tr;
    int stack[RL_STACK_SIZE];
    if (sv.n_groups) { // flat group table: root test, then scan + exact tests
        V3 inv = V3{1.0f / d.x, 1.0f / d.y, 1.0f / d.z};
        if (!aabb_intersect_ref(sv.root_min, sv.root_max, o, inv, RL_EPSILON, RL_F32_MAX)) {
            HitRec h;
            h.t = RL_F32_MAX, h.u = 0.0f, h.v = 0.0f, h.prim = RL_MISS;
            return h;
        }
        return flat_closest(sv, flat, trav, o, d);
    }
    if (closest_begin(tr, sv, o, d)) {
#if RL_TREE_DIST
        float sdist[RL_STACK_SIZE];
#else
        float *sdist = nullptr;
#endif
        while (tr.cur != RL_TRAV_DONE) {
            if (tr.cur >= 0) trav_node_step<false>(tr, stack, nodes, sdist);
            else trav_leaf_closest(tr, stack, trav, sdist);
        }
    }
    HitRec h = closest_result(tr);
    if (RL_REF_ORDER && sv.ref_nodes && closest_ambiguous(tr, sv)) ref_bvh_closest(sv, trav, o, d, &h.t, &h.u, &h.v, &h.prim);
    return h;
}
RL_HD bool trace_visible(const SceneView &sv, const float4 *flat, const float4 *nodes, const float4 *trav, V3 p0, V3 p1) {
    Trav tr;
    int stack[RL_STACK_SIZE];
    bool decided, vis;
    if (sv.n_groups) {
        V3 d;
        float thr;
        if (!visible_setup(sv, p0, p1, &d, &thr)) return false; // accel.rs:338-340
        return !flat_any(sv, flat, trav, p0, d, thr);
    }
    visible_begin(tr, sv, p0, p1, &decided, &vis);
    if (decided) return vis;
    while (tr.cur != RL_TRAV_DONE) {
        if (tr.cur >= 0) trav_node_step<true>(tr, stack, nodes);
        else if (trav_leaf_any(tr, stack, trav)) return false;
    }
    if (tr.amb) return !(ref_path_ok(sv, tr.rim_slot, tr.o, tr.d, tr.tmax) || ref_bvh_any(sv, trav, tr.o, tr.d, tr.tmax)); // blocked by rim hits only
    return true;
}
// the group table where the scene description says it is (global memory); kernels that stage it pass their copy
RL_HD HitRec trace_closest(const SceneView &sv, const float4 *nodes, const float4 *trav, V3 o, V3 d) { return trace_closest(sv, sv.flat, nodes, trav, o, d); }
RL_HD bool trace_visible(const SceneView &sv, const float4 *nodes, const float4 *trav, V3 p0, V3 p1) { return trace_visible(sv, sv.flat, nodes, trav, p0, p1); }

// ---- Camera::generate (camera.rs:81-91) -------------------------------------------------------
RL_HD void m4_mul_v4(const float *m, float x, float y, float z, float w, float *out) {
    for (int r = 0; r < 4; r++) out[r] = ((m[r] * x + m[4 + r] * y) + m[8 + r] * z) + m[12 + r] * w;
}
RL_HD void camera_generate(const SceneView &sv, float px, float py, V3 *o, V3 *d) {
    float h[4];
    m4_mul_v4(sv.s2c, px / sv.img_w, py / sv.img_h, 0.0f, 1.0f, h);
    float iw = 1.0f / h[3];
    V3 near_p = V3{h[0] * iw, h[1] * iw, h[2] * iw};
    V3 dl = normalize(near_p);
    float g[4];
    m4_mul_v4(sv.c2w, dl.x, dl.y, dl.z, 0.0f, g);
    *o = sv.cam_pos;
    *d = V3{g[0], g[1], g[2]};
}

// ---- camera rays of group-table scenes: which quads can the pixels [x0, x1] x [y0, y1] see? ---------------------
// The pixel-space rectangle (jitter included, grown by 0.05 px), the pyramid of its four corner rays through the camera
// position (Camera::generate in double), and per quad of the table the classic conservative frustum test: a quad whose
// vertices (A and B, quad_verts[6 bit ..]) all lie more than eps outside ONE side plane cannot contain a point of any ray of
// the block.  eps = 1e-4 x (scene + camera extent) dominates the rounding of the float ray directions (~1e-7 t) and of the
// exact test's own acceptance region (ulps of the barycentrics).
RL_HD uint32_t camera_block_mask(const SceneView &sv, const float4 *quad_verts, uint32_t valid_quads, uint32_t x0, uint32_t x1, uint32_t y0, uint32_t y1,
                                 double eps) {
    const double rx[2] = {(double)x0 - 0.05, (double)x1 + 1.05}, ry[2] = {(double)y0 - 0.05, (double)y1 + 1.05};
    double D[4][3];
    for (int c = 0; c < 4; c++) { // corners in cyclic order
        const double px = rx[(c == 1 || c == 2) ? 1 : 0] / (double)sv.img_w, py = ry[c >= 2 ? 1 : 0] / (double)sv.img_h;
        double h[4], g[3];
        for (int r = 0; r < 4; r++) h[r] = (double)sv.s2c[r] * px + (double)sv.s2c[4 + r] * py + (double)sv.s2c[12 + r];
        for (int r = 0; r < 3; r++) h[r] /= h[3];
        for (int r = 0; r < 3; r++) g[r] = (double)sv.c2w[r] * h[0] + (double)sv.c2w[4 + r] * h[1] + (double)sv.c2w[8 + r] * h[2];
        const double l = sqrt(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
        for (int r = 0; r < 3; r++) D[c][r] = g[r] / l;
    }
    const double ctr[3] = {D[0][0] + D[1][0] + D[2][0] + D[3][0], D[0][1] + D[1][1] + D[2][1] + D[3][1], D[0][2] + D[1][2] + D[2][2] + D[3][2]};
    double N[4][3];
    for (int c = 0; c < 4; c++) {
        const double *a = D[c], *e = D[(c + 1) & 3];
        double nx = a[1] * e[2] - a[2] * e[1], ny = a[2] * e[0] - a[0] * e[2], nz = a[0] * e[1] - a[1] * e[0];
        const double l = sqrt(nx * nx + ny * ny + nz * nz);
        if (!(l > 1e-12)) return valid_quads; // degenerate pyramid (or NaN): see everything
        const double sgn = (nx * ctr[0] + ny * ctr[1] + nz * ctr[2]) < 0.0 ? -1.0 : 1.0;
        N[c][0] = sgn * nx / l, N[c][1] = sgn * ny / l, N[c][2] = sgn * nz / l;
    }
    uint32_t m = valid_quads;
    for (uint32_t bit = 0; bit < 32u; bit++) {
        if (!((valid_quads >> bit) & 1u)) continue;
        bool culled = false;
        for (int c = 0; c < 4 && !culled; c++) {
            bool all_out = true;
            for (int k = 0; k < 6; k++) {
                const float4 v = quad_verts[6u * bit + k];
                const double dist = N[c][0] * ((double)v.x - (double)sv.cam_pos.x) + N[c][1] * ((double)v.y - (double)sv.cam_pos.y) +
                                    N[c][2] * ((double)v.z - (double)sv.cam_pos.z);
                if (!(dist < -eps)) all_out = false; // NaN vertices keep the quad
            }
            culled = all_out;
        }
        if (culled) m &= ~(1u << bit);
    }
    return m;
}

// ---- materials ---------------------------------------------------------------------------------
#define RL_MAT_F4 6
struct Material {
    Col kd, ks, le;       // kd: diffuse | metal eta | glass transmittance;  ks: specular / reflectance
    float exponent;       // phong exponent | microfacet alpha
    float weight_specular; // phong lobe weight | glass eta
    float inv_area, pdf_sel;
    uint32_t kind, microfacet;
    bool is_light;
    uint32_t emit_kind;   // rl_emission_kind; HSV / TEXTURE: le = {scale, texture index (bits), -} until mesh_emit has replaced it (apply_textures, sample_light)
    bool k_textured;      // metal k comes from a texture (k_tex) instead of ext[0]
    Col k_tex;
    const float4 *ext;    // row 4 {metal k.rgb, glass 1/eta}, row 5 {textures of the colour slots}: read only by the metal / glass branches and the texture lookup
};
RL_HD Material load_material(const float4 *mats, uint32_t mesh) {
    const float4 *row = mats + RL_MAT_F4 * mesh;
    float4 a = row[0], b = row[1], c = row[2], e = row[3];
    Material m;
    m.kd = xyz_col(a);
    m.kind = f2u(a.w);
    m.ks = xyz_col(b);
    m.exponent = b.w;
    m.le = xyz_col(c);
    m.emit_kind = f2u(c.w);
    m.is_light = m.emit_kind != 0u;
    m.weight_specular = e.x;
    m.inv_area = e.y;
    m.pdf_sel = e.z;
    m.microfacet = f2u(e.w);
    m.ext = row + 4;
    m.k_textured = false;
    m.k_tex = Col{0.0f, 0.0f, 0.0f};
    return m;
}
RL_HD Col metal_k(const Material &m) { return m.k_textured ? m.k_tex : xyz_col(m.ext[0]); } // BSDFMetal.k: constant or texture (apply_textures)
// bsdf_type().is_smooth() (bsdfs/mod.rs:157-161): DELTA in the type -> no light sampling, no MIS at this vertex
// KM (here and below): compile-time mask of the rl_bsdf_kind values present in the scene (bit k = kind k).  The shade
// kernel is instantiated for the masks {diffuse}, {diffuse, phong} and "all", so that a Cornell box does not carry the
// microfacet / Fresnel code (registers, instruction cache) it never runs; every other caller uses the default "all".
#define RL_KM_ALL 0x1ffu  // bits 0-7: BSDF kinds (bit 5: BSDFBlend), bit 8: the scene has textures
#define RL_HAS(KM, k) (((KM) >> (k)) & 1u)
template <uint32_t KM = RL_KM_ALL>
RL_HD bool mat_is_smooth(const Material &m) {
    return (RL_HAS(KM, 3) && m.kind == 3u) || (((RL_HAS(KM, 2) && m.kind == 2u) || (RL_HAS(KM, 4) && m.kind == 4u)) && m.microfacet == 0u);
}
template <uint32_t KM = RL_KM_ALL>
RL_HD bool mat_is_twosided(const Material &m) { return !(RL_HAS(KM, 3) && m.kind == 3u); } // glass.rs:181-183
RL_HD V3 reflect_local(V3 d) { return V3{-d.x, -d.y, d.z}; }

// ---- bsdfs/utils.rs -------------------------------------------------------------------------------
RL_HD Col operator-(Col a, Col b) { return Col{a.r - b.r, a.g - b.g, a.b - b.b}; }     // structure.rs:349-358
RL_HD Col col_div(Col a, Col b) { return Col{a.r / b.r, a.g / b.g, a.b / b.b}; }       // Div<Color>, :266-275
RL_HD Col col_value(float v) { return Col{v, v, v}; }
RL_HD Col col_safe_sqrt(Col c) { return Col{sqrtf(fmaxf(c.r, 0.0f)), sqrtf(fmaxf(c.g, 0.0f)), sqrtf(fmaxf(c.b, 0.0f))}; } // :132-138
RL_HD float powi2(float x) { return x * x; }
RL_HD float powi3(float x) { return x * (x * x); }
RL_HD float powi5(float x) { // llvm.powi / __powisf2: binary exponentiation, x * ((x^2)^2)
    float x2 = x * x;
    return x * (x2 * x2);
}
RL_HD float sin_2_theta(V3 w) { return fmaxf(1.0f - w.z * w.z, 0.0f); }   // utils.rs:14-16
RL_HD float tan_theta(V3 w) { return sqrtf(sin_2_theta(w)) / w.z; }       // :20-22
RL_HD float hypot2(float a, float b) {                                    // :50-60
    if (fabsf(a) > fabsf(b)) {
        float r = b / a;
        return fabsf(a) * sqrtf(1.0f + r * r);
    } else if (b != 0.0f) {
        float r = a / b;
        return fabsf(b) * sqrtf(1.0f + r * r);
    }
    return 0.0f;
}
RL_HD V3 reflect_vector(V3 wo, V3 n) { return -(wo) + n * 2.0f * dot(wo, n); } // :62-64
RL_HD bool check_reflection_condition(V3 wi, V3 wo) { return fabsf(wi.z * wo.z - wi.x * wo.x - wi.y * wo.y - 1.0f) < 0.0001f; } // :65-67
// fresnel_conductor, utils.rs:78-100
RL_HD Col fresnel_conductor(float cos_theta, Col eta, Col k) {
    float cos_theta_2 = cos_theta * cos_theta;
    float sin_theta_2 = 1.0f - cos_theta_2;
    float sin_theta_4 = sin_theta_2 * sin_theta_2;
    Col temp1 = eta * eta - k * k - col_value(sin_theta_2);
    Col a2pb2 = col_safe_sqrt(temp1 * temp1 + mul_checked(k * k * eta * eta, 4.0f));
    Col a = col_safe_sqrt(mul_checked(a2pb2 + temp1, 0.5f));
    Col term1 = a2pb2 + col_value(cos_theta_2);
    Col term2 = mul_checked(a, 2.0f * cos_theta_2);
    Col rs2 = col_div(term1 - term2, term1 + term2);
    Col term3 = mul_checked(a2pb2, cos_theta_2) + col_value(sin_theta_4);
    Col term4 = mul_checked(term2, sin_theta_2);
    Col rp2 = col_div(rs2 * (term3 - term4), term3 + term4);
    return mul_plain(0.5f, rp2 + rs2);
}
// fresnel_dielectric, utils.rs:103-130: (fresnel, cos_theta_t)
RL_HD void fresnel_dielectric(float cos_theta_i_, float eta, float *fresnel, float *cos_theta_t_out) {
    if (eta == 1.0f) {
        *fresnel = 0.0f;
        *cos_theta_t_out = -cos_theta_i_;
        return;
    }
    float scale = cos_theta_i_ > 0.0f ? 1.0f / eta : eta;
    float cos_theta_t_sqr = 1.0f - (1.0f - cos_theta_i_ * cos_theta_i_) * (scale * scale);
    if (cos_theta_t_sqr <= 0.0f) {
        *fresnel = 1.0f;
        *cos_theta_t_out = 0.0f;
        return;
    }
    float cos_theta_i = fabsf(cos_theta_i_);
    float cos_theta_t = sqrtf(cos_theta_t_sqr);
    float rs = (cos_theta_i - eta * cos_theta_t) / (cos_theta_i + eta * cos_theta_t);
    float rp = (eta * cos_theta_i - cos_theta_t) / (eta * cos_theta_i + cos_theta_t);
    *cos_theta_t_out = cos_theta_i_ > 0.0f ? -cos_theta_t : cos_theta_t;
    *fresnel = 0.5f * (rs * rs + rp * rp);
}

// ---- MicrofacetDistribution (bsdfs/distribution.rs:26-144), isotropic: alpha_u == alpha_v == alpha -----
// kind: 1 GGX, 2 Beckmann (rl_microfacet)
RL_HD float mf_eval(uint32_t kind, float alpha, V3 m) { // :27-56
    if (m.z <= 0.0f) return 0.0f;
    float cos_theta_2 = m.z * m.z;
    float beckmann_exp = ((m.x * m.x) / (alpha * alpha) + (m.y * m.y) / (alpha * alpha)) / cos_theta_2;
    float res;
    if (kind == 2u) res = spec_expf(-beckmann_exp) / (RL_PI * alpha * alpha * cos_theta_2 * cos_theta_2);
    else {
        float root = (1.0f + beckmann_exp) * cos_theta_2;
        res = 1.0f / (RL_PI * alpha * alpha * root * root);
    }
    return res * m.z < 1e-20f ? 0.0f : res;
}
RL_HD float mf_pdf(uint32_t kind, float alpha, V3 m) { return mf_eval(kind, alpha, m) * m.z; } // :58-60
RL_HD V3 mf_sample(uint32_t kind, float alpha, float sx, float sy, float *pdf_out) { // :63-111
    float sin_phi_m, cos_phi_m;
    spec_sincos(2.0f * RL_PI * sy, &sin_phi_m, &cos_phi_m);
    float alpha_sqr = alpha * alpha;
    float cos_theta_m, pdf;
    if (kind == 2u) {
        float tan_theta_m_sqr = alpha_sqr * -spec_logf(1.0f - sx);
        cos_theta_m = 1.0f / sqrtf(1.0f + tan_theta_m_sqr);
        pdf = (1.0f - sx) / (RL_PI * alpha * alpha * powi3(cos_theta_m));
    } else {
        float tan_theta_m_sqr = alpha_sqr * sx / (1.0f - sx);
        cos_theta_m = 1.0f / sqrtf(1.0f + tan_theta_m_sqr);
        float tmp = 1.0f + tan_theta_m_sqr / alpha_sqr;
        pdf = RL_FRAC_1_PI / (alpha * alpha * powi3(cos_theta_m) * powi2(tmp));
    }
    if (pdf < 1e-20f) pdf = 0.0f;
    float sin_theta_m = sqrtf(fmaxf(1.0f - powi2(cos_theta_m), 0.0f));
    *pdf_out = pdf;
    return V3{sin_theta_m * cos_phi_m, sin_theta_m * sin_phi_m, cos_theta_m};
}
RL_HD float mf_smith_g1(uint32_t kind, float alpha, V3 v, V3 m) { // :117-144
    if (dot(v, m) * v.z <= 0.0f) return 0.0f;
    float tan_t = fabsf(tan_theta(v));
    if (tan_t == 0.0f) return 1.0f;
    if (kind == 2u) {
        float a = 1.0f / (alpha * tan_t);
        if (a >= 1.6f) return 1.0f;
        float a_sqr = powi2(a);
        return (3.535f * a + 2.181f * a_sqr) / (1.0f + 2.276f * a + 2.577f * a_sqr);
    }
    float root = alpha * tan_t;
    return 2.0f / (1.0f + hypot2(1.0f, root));
}
RL_HD float mf_g(uint32_t kind, float alpha, V3 wi, V3 wo, V3 m) { return mf_smith_g1(kind, alpha, wi, m) * mf_smith_g1(kind, alpha, wo, m); } // :113-115

// ---- BSDFMetal with a distribution (bsdfs/metal.rs:37-70, 95-109, 128-155): GLOSSY, SolidAngle pdfs ------
RL_HD float metal_pdf(const Material &mt, V3 wi, V3 wo) {
    V3 h = normalize(wi + wo);
    return mf_pdf(mt.microfacet, mt.exponent, h) / (4.0f * fabsf(dot(wo, h)));
}
RL_HD Col metal_eval(const Material &mt, V3 wi, V3 wo) {
    V3 h = normalize(wi + wo);
    float d = mf_eval(mt.microfacet, mt.exponent, h);
    if (d == 0.0f) return Col{0.0f, 0.0f, 0.0f};
    Col f = mt.ks * fresnel_conductor(dot(wi, h), mt.kd, metal_k(mt));
    float g = mf_g(mt.microfacet, mt.exponent, wi, wo, h);
    float model = d * g / (4.0f * wi.z);
    return mul_checked(f, model);
}
// ---- BSDFSubstrate (bsdfs/substrate.rs): pdf / eval per domain ---------------------------------------
RL_HD Col substrate_schlick(const Material &mt, float cos_theta) { // :15-18
    Col rs = mt.ks;
    return rs + mul_checked(col_value(1.0f) - rs, powi5(1.0f - cos_theta));
}
RL_HD float substrate_pdf(const Material &mt, V3 wi, V3 wo, bool discrete) { // :92-147
    if (wi.z <= 0.0f || wo.z <= 0.0f) return 0.0f;
    V3 m = wi + wo;
    if (m.x == 0.0f && m.y == 0.0f && m.z == 0.0f) return 0.0f;
    m = normalize(m);
    if (discrete) return 0.5f; // reached only with wo = reflect(wi): check_reflection_condition holds
    float pdf_diffuse = wo.z * RL_FRAC_1_PI;
    float pdf_specular = mt.microfacet == 0u ? 0.0f : mf_pdf(mt.microfacet, mt.exponent, m) / (4.0f * fabsf(dot(wo, m)));
    return 0.5f * (pdf_diffuse + pdf_specular);
}
RL_HD Col substrate_eval(const Material &mt, V3 d_in, V3 d_out, bool discrete) { // :149-206
    if (d_in.z <= 0.0f || d_out.z <= 0.0f) return Col{0.0f, 0.0f, 0.0f};
    V3 m = d_in + d_out;
    if (m.x == 0.0f && m.y == 0.0f && m.z == 0.0f) return Col{0.0f, 0.0f, 0.0f};
    m = normalize(m);
    if (discrete) return substrate_schlick(mt, dot(d_in, m));
    Col diffuse = mul_checked(mul_checked(mul_checked(mt.kd * (col_value(1.0f) - mt.ks), 28.0f / (23.0f * RL_PI)),
                                          1.0f - powi5(1.0f - 0.5f * fabsf(d_in.z))),
                              1.0f - powi5(1.0f - 0.5f * fabsf(d_out.z)));
    Col specular = Col{0.0f, 0.0f, 0.0f};
    if (mt.microfacet != 0u) {
        float model = mf_eval(mt.microfacet, mt.exponent, m) / (4.0f * fabsf(dot(d_in, m)) * fmaxf(fabsf(d_in.z), fabsf(d_out.z)));
        specular = mul_plain(model, substrate_schlick(mt, dot(d_in, m)));
    }
    return mul_checked(diffuse + specular, d_out.z);
}

// BSDF::pdf (diffuse.rs:33-51, phong.rs:65-91)
template <uint32_t KM = RL_KM_ALL>
RL_HD float bsdf_pdf_leaf(const Material &m, V3 wi, V3 wo) {
    if (RL_HAS(KM, 2) && m.kind == 2u) return metal_pdf(m, wi, wo);               // only reached with a distribution (not smooth)
    if (RL_HAS(KM, 4) && m.kind == 4u) return substrate_pdf(m, wi, wo, false);
    if (!RL_HAS(KM, 1) || m.kind == 0u) {
        if (wi.z <= 0.0f) return 0.0f;
        if (wo.z <= 0.0f) return 0.0f;
        return wo.z * RL_FRAC_1_PI;
    }
    if (wi.z <= 0.0f || wo.z <= 0.0f) return 0.0f;
    float pdf_specular;
    float alpha = dot(reflect_local(wi), wo);
    if (alpha > 0.0f) pdf_specular = m.weight_specular * spec_powf(alpha, m.exponent) * (m.exponent + 1.0f) / (2.0f * RL_PI);
    else pdf_specular = 0.0f;
    float pdf_diffuse = (1.0f - m.weight_specular) * wo.z * RL_FRAC_1_PI;
    return pdf_specular + pdf_diffuse;
}
// BSDF::eval (diffuse.rs:53-71, phong.rs:93-119); includes the cosine for the diffuse lobe
template <uint32_t KM = RL_KM_ALL>
RL_HD Col bsdf_eval_leaf(const Material &m, V3 wi, V3 wo) {
    if (RL_HAS(KM, 2) && m.kind == 2u) return metal_eval(m, wi, wo);
    if (RL_HAS(KM, 4) && m.kind == 4u) return substrate_eval(m, wi, wo, false);
    if (!RL_HAS(KM, 1) || m.kind == 0u) {
        if (wi.z <= 0.0f) return Col{0.0f, 0.0f, 0.0f};
        if (wo.z > 0.0f) return mul_checked(mul_checked(m.kd, wo.z), RL_FRAC_1_PI);
        return Col{0.0f, 0.0f, 0.0f};
    }
    if (wi.z <= 0.0f || wo.z <= 0.0f) return Col{0.0f, 0.0f, 0.0f};
    Col specular_value;
    float alpha = dot(reflect_local(wi), wo);
    if (alpha > 0.0f) specular_value = mul_checked(m.ks, spec_powf(alpha, m.exponent) * (m.exponent + 2.0f) / (2.0f * RL_PI));
    else specular_value = Col{0.0f, 0.0f, 0.0f};
    Col diffuse_value = mul_checked(mul_checked(m.kd, wo.z), RL_FRAC_1_PI);
    return specular_value + diffuse_value;
}
// Phong pdf and eval of one direction pair with ONE evaluation of alpha^n (phong.rs:65-119): identical bits to calling
// bsdf_pdf and bsdf_eval separately, half the transcendental work
RL_HD void phong_eval_pdf(const Material &m, V3 wi, V3 wo, Col *f, float *pdf) {
    if (wi.z <= 0.0f || wo.z <= 0.0f) {
        *f = Col{0.0f, 0.0f, 0.0f};
        *pdf = 0.0f;
        return;
    }
    float alpha = dot(reflect_local(wi), wo);
    float pdf_specular = 0.0f;
    Col specular_value = Col{0.0f, 0.0f, 0.0f};
    if (alpha > 0.0f) {
        const float p = spec_powf(alpha, m.exponent);
        pdf_specular = m.weight_specular * p * (m.exponent + 1.0f) / (2.0f * RL_PI);
        specular_value = mul_checked(m.ks, p * (m.exponent + 2.0f) / (2.0f * RL_PI));
    }
    float pdf_diffuse = (1.0f - m.weight_specular) * wo.z * RL_FRAC_1_PI;
    *pdf = pdf_specular + pdf_diffuse;
    *f = specular_value + mul_checked(mul_checked(m.kd, wo.z), RL_FRAC_1_PI);
}
// eval and pdf of one direction pair (what light sampling with MIS needs)
template <uint32_t KM = RL_KM_ALL>
RL_HD void bsdf_eval_pdf_leaf(const Material &m, V3 wi, V3 wo, Col *f, float *pdf) {
    if (RL_HAS(KM, 1) && m.kind == 1u) {
        phong_eval_pdf(m, wi, wo, f, pdf);
        return;
    }
    *f = bsdf_eval_leaf<KM>(m, wi, wo);
    *pdf = bsdf_pdf_leaf<KM>(m, wi, wo);
}
// BSDFBlend (bsdfs/blend.rs): kind 5, m.weight_specular = weight, ext[0].xy = row distance to the two parts (rl_scene_host.hpp: material_rows)
#define RL_KM_NOBLEND(KM) ((KM) & ~(1u << 5))
RL_HD Material blend_part(const Material &m, uint32_t which) {
    const float4 e = m.ext[0];
    return load_material(m.ext - 4 + (int32_t)f2u(which == 0u ? e.x : e.y), 0u);
}
// eval: weight * bsdf1.eval + (1 - weight) * bsdf2.eval (`f32 * Color`, structure.rs:294-303: plain products); pdf: pdf_1 * weight + pdf_2 * (1 - weight)
template <uint32_t KM = RL_KM_ALL>
RL_HD void bsdf_eval_pdf(const Material &m, V3 wi, V3 wo, Col *f, float *pdf) {
    if (RL_HAS(KM, 5) && m.kind == 5u) {
        const Material a = blend_part(m, 0u), b = blend_part(m, 1u);
        Col fa, fb;
        float pa, pb;
        bsdf_eval_pdf_leaf<RL_KM_NOBLEND(KM)>(a, wi, wo, &fa, &pa);
        bsdf_eval_pdf_leaf<RL_KM_NOBLEND(KM)>(b, wi, wo, &fb, &pb);
        const float w = m.weight_specular;
        *pdf = pa * w + pb * (1.0f - w);
        *f = mul_plain(w, fa) + mul_plain(1.0f - w, fb);
        return;
    }
    bsdf_eval_pdf_leaf<KM>(m, wi, wo, f, pdf);
}
template <uint32_t KM = RL_KM_ALL>
RL_HD float bsdf_pdf(const Material &m, V3 wi, V3 wo) {
    if (RL_HAS(KM, 5) && m.kind == 5u) {
        Col f;
        float p;
        bsdf_eval_pdf<KM>(m, wi, wo, &f, &p);
        return p;
    }
    return bsdf_pdf_leaf<KM>(m, wi, wo);
}
template <uint32_t KM = RL_KM_ALL>
RL_HD Col bsdf_eval(const Material &m, V3 wi, V3 wo) {
    if (RL_HAS(KM, 5) && m.kind == 5u) {
        Col f;
        float p;
        bsdf_eval_pdf<KM>(m, wi, wo, &f, &p);
        return f;
    }
    return bsdf_eval_leaf<KM>(m, wi, wo);
}
// BSDF::sample (diffuse.rs:11-31, phong.rs:14-63, metal.rs:15-73, glass.rs:75-121, substrate.rs:22-90).
// *discrete: the sampled pdf is PDF::Discrete (delta lobe) -- no MIS for the edge it creates.
template <uint32_t KM = RL_KM_ALL>
RL_HD bool bsdf_sample_leaf(const Material &m, V3 wi, float sx, float sy, Col *weight, V3 *wo, float *pdf, bool *discrete) {
    *discrete = false;
    if (RL_HAS(KM, 3) && m.kind == 3u) { // glass: no wi.z test (not two-sided), transport == Importance -> factor 1
        float fresnel, cos_theta_trans;
        fresnel_dielectric(wi.z, m.weight_specular, &fresnel, &cos_theta_trans);
        *discrete = true;
        *pdf = fresnel;
        if (sx <= fresnel) {
            *weight = m.ks;
            *wo = reflect_local(wi);
        } else {
            *weight = mul_checked(mul_checked(m.kd, 1.0f), 1.0f);
            float scale = cos_theta_trans < 0.0f ? -m.ext[0].w : -m.weight_specular;
            *wo = V3{scale * wi.x, scale * wi.y, cos_theta_trans};
        }
        return true;
    }
    if (wi.z <= 0.0f) return false;
    if (RL_HAS(KM, 2) && m.kind == 2u) { // metal
        Col k = metal_k(m);
        if (m.microfacet == 0u) {
            *weight = m.ks * fresnel_conductor(wi.z, m.kd, k);
            *wo = reflect_local(wi);
            *pdf = 1.0f;
            *discrete = true;
            return true;
        }
        float p;
        V3 mm = mf_sample(m.microfacet, m.exponent, sx, sy, &p);
        if (p == 0.0f) return false;
        V3 d_out = reflect_vector(wi, mm);
        if (d_out.z <= 0.0f) return false;
        Col f = fresnel_conductor(dot(wi, mm), m.kd, k) * m.ks;
        float w = mf_eval(m.microfacet, m.exponent, mm) * mf_g(m.microfacet, m.exponent, wi, d_out, mm) * dot(wi, mm) / (p * wi.z);
        *weight = mul_plain(w, f);
        *wo = d_out;
        *pdf = p; // the microfacet-normal pdf, as the reference returns it (metal.rs:64)
        return true;
    }
    if (RL_HAS(KM, 4) && m.kind == 4u) { // substrate
        V3 d_out;
        bool disc = false;
        if (sx < 0.5f) {
            sx = sx * 2.0f;
            d_out = cosine_sample_hemisphere(sx, sy);
        } else {
            sx = (sx - 0.5f) * 2.0f;
            V3 mm;
            if (m.microfacet == 0u) {
                mm = V3{0.0f, 0.0f, 1.0f};
                disc = true;
            } else {
                float p;
                mm = mf_sample(m.microfacet, m.exponent, sx, sy, &p);
                if (p == 0.0f) return false;
            }
            d_out = reflect_vector(wi, mm);
            if (d_out.z <= 0.0f) return false;
        }
        float p = substrate_pdf(m, wi, d_out, disc);
        if (p == 0.0f) return false;
        *weight = div_checked(substrate_eval(m, wi, d_out, disc), p);
        *wo = d_out;
        *pdf = p;
        *discrete = disc;
        return true;
    }
    if (!RL_HAS(KM, 1) || m.kind == 0u) {
        V3 d_out = cosine_sample_hemisphere(sx, sy);
        *weight = m.kd;
        *wo = d_out;
        *pdf = d_out.z * RL_FRAC_1_PI;
        return true;
    }
    V3 d_out;
    if (sx < m.weight_specular) {
        sx = sx / m.weight_specular;
        const double log2_sy = (sy > 0.0f && sy != 1.0f) ? spec_log2((double)sy) : 0.0;
        float sin_alpha = sqrtf(1.0f - spec_powf_pre(sy, 2.0f / (m.exponent + 1.0f), log2_sy));
        float cos_alpha = spec_powf_pre(sy, 1.0f / (m.exponent + 1.0f), log2_sy);
        float phi = 2.0f * RL_PI * sx;
        float sp, cp;
        spec_sincos(phi, &sp, &cp);
        V3 local_dir = V3{sin_alpha * cp, sin_alpha * sp, cos_alpha};
        Frame fr = make_frame(reflect_local(wi));
        d_out = to_world(fr, local_dir);
        if (d_out.z <= 0.0f) return false;
    } else {
        sx = (sx - m.weight_specular) / (1.0f - m.weight_specular);
        d_out = cosine_sample_hemisphere(sx, sy);
    }
    float p;
    Col f;
    phong_eval_pdf(m, wi, d_out, &f, &p);
    if (p == 0.0f) return false;
    *weight = div_checked(f, p);
    *wo = d_out;
    *pdf = p;
    return true;
}
template <uint32_t KM = RL_KM_ALL>
RL_HD bool bsdf_sample(const Material &m, V3 wi, float sx, float sy, Col *weight, V3 *wo, float *pdf, bool *discrete) {
    if (RL_HAS(KM, 5) && m.kind == 5u) { // BSDFBlend::sample, blend.rs:10-45: pick a part by sample.x, then pdf and eval of the whole blend
        const float w = m.weight_specular;
        const bool first = sx < w;
        const Material part = blend_part(m, first ? 0u : 1u);
        const float sx2 = first ? sx * (1.0f / w) : (sx - w) * (1.0f / (1.0f - w));
        Col w_part;
        float p_part;
        V3 d_out;
        if (!bsdf_sample_leaf<RL_KM_NOBLEND(KM)>(part, wi, sx2, sy, &w_part, &d_out, &p_part, discrete)) return false;
        Col f;
        float p;
        bsdf_eval_pdf<KM>(m, wi, d_out, &f, &p);
        if (p == 0.0f) return false;
        *weight = div_checked(f, p);
        *wo = d_out;
        *pdf = p;
        *discrete = false;
        return true;
    }
    return bsdf_sample_leaf<KM>(m, wi, sx, sy, weight, wo, pdf, discrete);
}

// ---- BSDFColor::color for the kd slot (bsdfs/mod.rs:31-101) -----------------------------------------
RL_HD uint64_t f32_as_usize(float v) { // Rust `as usize`: saturating, NaN -> 0
    if (!(v > 0.0f)) return 0ull;
    if (v >= 18446744073709551616.0f) return ~0ull;
    return (uint64_t)v;
}
RL_HD int32_t f32_as_i32(float v) { // Rust `as i32`: saturating, NaN -> 0
    if (v != v) return 0;
    if (v >= 2147483648.0f) return 2147483647;
    if (v <= -2147483648.0f) return (int32_t)0x80000000u;
    return (int32_t)v;
}
RL_HD float modulo1(float x) { return fmodf(fmodf(x, 1.0f) + 1.0f, 1.0f); } // tools.rs:39-41 with n = 1.0
// Bitmap::pixel_uv (structure.rs:434-453) of bitmap texture tex_id
RL_HD Col bitmap_pixel_uv(const SceneView &sv, uint32_t tex_id, float ux, float uy) {
    const float4 t3 = sv.tex[4 * tex_id + 3];
    const uint32_t width = f2u(t3.x), height = f2u(t3.y), off = f2u(t3.z);
    ux = modulo1(ux), uy = modulo1(uy);
    const uint64_t x = f32_as_usize(ux * (float)width), y = f32_as_usize(uy * (float)height);
    const uint64_t i = (uint64_t)width * y + x;
    if (i >= (uint64_t)width * height) return Col{0.0f, 0.0f, 0.0f};
    return xyz_col(sv.texels[off + i]);
}
// Mesh::emit(uv) for EmissionType::HSV / Texture (geometry.rs:184-206): `m.le` still holds {scale, texture index}.
// HSV: c = x * (1, 0, 0) + (1 - x) * (0, 1, 0) with x = uv.x.abs() % 1.0 (`f32 * Color`: plain products), then Color * scale.
RL_HD Col mesh_emit(const SceneView &sv, const Material &m, float ux, float uy) {
    const float scale = m.le.r;
    Col c;
    if (m.emit_kind == 2u) {
        const float x = fmodf(fabsf(ux), 1.0f);
        c = Col{x * 1.0f + (1.0f - x) * 0.0f, x * 0.0f + (1.0f - x) * 1.0f, x * 0.0f + (1.0f - x) * 0.0f};
    } else c = bitmap_pixel_uv(sv, f2u(m.le.g), ux, uy);
    return mul_checked(c, scale);
}
// kd of a textured material at a hit: `flags` bit 1 = the mesh has uv; (hit_u, hit_v) barycentrics of the hit
RL_HD Col texture_kd(const SceneView &sv, uint32_t tex_id, uint32_t prim, uint32_t flags, float hit_u, float hit_v) {
    if (!(flags & 2u)) return Col{0.0f, 0.0f, 0.0f}; // "Found a texture but no uv coordinate given"
    // uv interpolation, structure.rs:1015-1023: d0 * (1 - u - v) + d1 * u + d2 * v
    const float2 d0 = sv.uvs[3 * prim], d1 = sv.uvs[3 * prim + 1], d2 = sv.uvs[3 * prim + 2];
    const float w = 1.0f - hit_u - hit_v;
    float ux = d0.x * w + d1.x * hit_u + d2.x * hit_v, uy = d0.y * w + d1.y * hit_u + d2.y * hit_v;
    const float4 t0 = sv.tex[4 * tex_id], t1 = sv.tex[4 * tex_id + 1], t2 = sv.tex[4 * tex_id + 2];
    const uint32_t kind = f2u(t0.w);
    if (kind == 1u) return bitmap_pixel_uv(sv, tex_id, ux, uy); // Bitmap::pixel_uv, structure.rs:434-453
    if (kind == 2u) { // Checkerbord, bsdfs/mod.rs:43-65
        ux = ux * t2.z + t2.x, uy = uy * t2.w + t2.y;
        const int32_t x = 2 * (f32_as_i32(ux * 2.0f) % 2) - 1, y = 2 * (f32_as_i32(uy * 2.0f) % 2) - 1;
        return x * y == 1 ? xyz_col(t0) : xyz_col(t1);
    }
    // Grid, bsdfs/mod.rs:66-99 (note `uv.y + scale.y`)
    ux = ux * t2.z + t2.x, uy = (uy + t2.w) + t2.y;
    float x = ux - floorf(ux), y = uy - floorf(uy);
    if (x > 0.5f) x -= 1.0f;
    if (y > 0.5f) y -= 1.0f;
    return (fabsf(x) < t1.w || fabsf(y) < t1.w) ? xyz_col(t0) : xyz_col(t1);
}
// BSDFColor::color(uv) of every textured colour slot of the material at this hit: slot a (kd | metal eta | glass kt) and slot b (ks)
// replace the constants loaded by load_material; slot c (metal k) goes to m.k_tex (the metal branches read k through metal_k()).
RL_HD void apply_textures(const SceneView &sv, Material &m, uint32_t prim, uint32_t flags, float hit_u, float hit_v) {
    const float4 t = m.ext[1];
    const uint32_t ta = f2u(t.x), tb = f2u(t.y), tc = f2u(t.z);
    if (ta != 0u) m.kd = texture_kd(sv, ta - 1u, prim, flags, hit_u, hit_v);
    if (tb != 0u) m.ks = texture_kd(sv, tb - 1u, prim, flags, hit_u, hit_v);
    if (tc != 0u) m.k_tex = texture_kd(sv, tc - 1u, prim, flags, hit_u, hit_v), m.k_textured = true;
    if (m.emit_kind >= 2u) { // its.mesh.emit(&its.uv): the uv of the hit (structure.rs:1015-1023); such meshes always carry uv (scene build)
        const float2 d0 = sv.uvs[3 * prim], d1 = sv.uvs[3 * prim + 1], d2 = sv.uvs[3 * prim + 2];
        const float w = 1.0f - hit_u - hit_v;
        m.le = mesh_emit(sv, m, d0.x * w + d1.x * hit_u + d2.x * hit_v, d0.y * w + d1.y * hit_u + d2.y * hit_v);
        m.emit_kind = 1u;
    }
}

// ---- surface interaction (fill_intersection, structure.rs:965-1059) ---------------------------
struct Surface {
    V3 p, n_g, n_s, wi;
    Frame frame;
    uint32_t mesh;
};
template <uint32_t KM = RL_KM_ALL>
RL_HD Surface fill_intersection(const SceneView &sv, const Material &mat, uint32_t prim, uint32_t mesh, float4 s0, float4 s1,
                                float4 s2, float4 s3, float t, float hit_u, float hit_v, V3 o, V3 d) {
    Surface s;
    s.mesh = mesh;
    s.p = o + t * d; // its.p as computed inside intersection_tri (geometry.rs:382)
    V3 n_g = xyz(s0);
    V3 n_s;
    if (f2u(s1.w) & 1u) {
        V3 d0 = xyz(s1), d1 = xyz(s2), d2 = xyz(s3);
        V3 ns = d0 * (1.0f - hit_u - hit_v) + d1 * hit_u + d2 * hit_v;
        if (dot(n_g, ns) < 0.0f) n_g = -n_g;
        float l = dot(ns, ns);
        if (l == 0.0f) n_s = n_g;
        else if (l != 1.0f) n_s = ns / sqrtf(l);
        else n_s = ns;
    } else n_s = n_g;
    // bsdf.is_twosided() && !is_light (structure.rs:1006): every BSDF but glass is two-sided; lights are never flipped
    if (mat_is_twosided<KM>(mat) && !mat.is_light && dot(d, n_s) > 0.0f) {
        n_s = V3{-n_s.x, -n_s.y, -n_s.z};
        n_g = V3{-n_g.x, -n_g.y, -n_g.z};
    }
    s.n_g = n_g;
    s.n_s = n_s;
    s.frame = make_frame(n_s);
    s.wi = to_local(s.frame, -d);
    (void)sv;
    (void)prim;
    return s;
}

// ---- emitters -----------------------------------------------------------------------------------
// Distribution1D::sample_discrete (math.rs:447-457): last index with cdf[i] <= v
RL_HD uint32_t cdf_sample_discrete(const float *cdf, uint32_t n_plus_1, float v) {
    uint32_t lo = 0, hi = n_plus_1; // first index with cdf[i] > v
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (cdf[mid] <= v) lo = mid + 1;
        else hi = mid;
    }
    return lo - 1;
}
// math.rs:324-352
RL_HD bool solve_quadratic(float a, float b, float c, float *x0_out, float *x1_out) {
    if (a == 0.0f) {
        if (b != 0.0f) {
            float v = -c / b;
            *x0_out = v, *x1_out = v;
            return true;
        }
        return false;
    }
    float d = b * b - 4.0f * a * c;
    if (d < 0.0f) return false;
    float d_sqrt = sqrtf(d);
    float tmp = b < 0.0f ? -0.5f * (b - d_sqrt) : -0.5f * (b + d_sqrt);
    float x0 = tmp / a, x1 = c / tmp;
    if (x0 > x1) *x0_out = x1, *x1_out = x0;
    else *x0_out = x0, *x1_out = x1;
    return true;
}
// BoundingSphere::intersect (structure.rs:899-920) for Ray::new(o, d) (tnear = EPSILON, tfar = f32::MAX).  Note the
// reference's b = +2 d_p.d with d_p = center - o (the roots are those of the mirrored ray); kept verbatim.
RL_HD bool bsphere_intersect(V3 center, float radius, V3 o, V3 d, float *t) {
    V3 d_p = center - o;
    float a = dot(d, d);
    float b = 2.0f * dot(d_p, d);
    float c = dot(d_p, d_p) - radius * radius;
    float t0, t1;
    if (!solve_quadratic(a, b, c, &t0, &t1)) return false;
    if (t0 < RL_EPSILON) {
        if (t1 < RL_F32_MAX) {
            *t = t1;
            return true;
        }
        return false;
    } else if (t0 < RL_F32_MAX) {
        *t = t0;
        return true;
    }
    return false;
}
RL_HD V3 sample_uniform_sphere(float ux, float uy) { // math.rs:67-72
    float z = 1.0f - 2.0f * ux;
    float r = sqrtf(fmaxf(1.0f - z * z, 0.0f));
    float phi = 2.0f * RL_PI * uy;
    float sp, cp;
    spec_sincos(phi, &sp, &cp);
    return V3{r * cp, r * sp, z};
}
#define RL_ENV_PDF (1.0f / (RL_PI * 4.0f)) // EnvironmentLightColor::Constant pdf (emitter.rs:406, 369-373)
// ---- EnvironmentLightColor::Texture (emitter.rs:300-427) over sv.env_dist / sv.texels ---------------------------------
RL_HD float clamp_ref(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); } // lib.rs:59-67 (NaN passes through)
#define RL_ONE_MINUS_EPSILON 0.9999999403953552f // lib.rs:52
RL_HD void to_spherical_coordinates(V3 d, float *u, float *v) { // emitter.rs:320-338
    float p = spec_atan2f(d.y, d.x);
    if (p < 0.0f) p = p + 2.0f * RL_PI;
    const float ux = p * RL_FRAC_1_PI * 0.5f;
    const float uy = spec_acosf(clamp_ref(d.z, -1.0f, 1.0f)) * RL_FRAC_1_PI;
    *u = clamp_ref(ux, 0.0f, RL_ONE_MINUS_EPSILON);
    *v = clamp_ref(uy, 0.0f, RL_ONE_MINUS_EPSILON);
}
RL_HD const float *env_cond_cdf(const SceneView &sv, uint64_t y) { return sv.env_dist + (sv.env_h + 1u) + y * (uint64_t)(sv.env_w + 1u); }
RL_HD float env_cdf_pdf(const SceneView &sv, uint64_t x, uint64_t y) { // Distribution2D::pdf, math.rs:529-531
    const float *func = sv.env_dist + (sv.env_h + 1u) + (uint64_t)sv.env_h * (sv.env_w + 1u);
    return func[y * sv.env_w + x] / sv.env_func_int;
}
RL_HD float dist1d_sample_continuous(const float *cdf, uint32_t n_plus_1, float v) { // math.rs:459-478
    const uint32_t i = cdf_sample_discrete(cdf, n_plus_1, v);
    float dv = v - cdf[i];
    const float pdf = cdf[i + 1] - cdf[i];
    if (pdf > 0.0f) dv = dv / pdf;
    return (float)i + dv;
}
RL_HD Col env_eval(const SceneView &sv, V3 d) { // EnvironmentLightColor::eval -> Bitmap::pixel_uv (structure.rs:434-453)
    float ux, uy;
    to_spherical_coordinates(d, &ux, &uy);
    ux = modulo1(ux), uy = modulo1(uy);
    const uint64_t x = f32_as_usize(ux * (float)sv.env_w), y = f32_as_usize(uy * (float)sv.env_h);
    const uint64_t i = (uint64_t)sv.env_w * y + x;
    if (i >= (uint64_t)sv.env_w * sv.env_h) return Col{0.0f, 0.0f, 0.0f};
    return xyz_col(sv.texels[sv.env_texel_off + i]);
}
RL_HD float env_pdf(const SceneView &sv, V3 d) { // EnvironmentLightColor::pdf, emitter.rs:403-425
    float ux, uy;
    to_spherical_coordinates(d, &ux, &uy);
    const float pdf = env_cdf_pdf(sv, f32_as_usize(ux * (float)sv.env_w), f32_as_usize(uy * (float)sv.env_h));
    float st, ct;
    spec_sincos(RL_PI * uy, &st, &ct);
    if (st == 0.0f) return 0.0f;
    return pdf / (2.0f * (RL_PI * RL_PI) * st);
}
RL_HD void env_sample_direction(const SceneView &sv, float sx, float sy, V3 *d, Col *color, float *pdf) { // emitter.rs:354-391
    float y = dist1d_sample_continuous(sv.env_dist, sv.env_h + 1u, sy); // Distribution2D::sample_continuous, math.rs:523-527
    float x = dist1d_sample_continuous(env_cond_cdf(sv, f32_as_usize(y)), sv.env_w + 1u, sx);
    x = clamp_ref(x, 0.0f, (float)sv.env_w - 1.0f);
    y = clamp_ref(y, 0.0f, (float)sv.env_h - 1.0f);
    const uint64_t xi = f32_as_usize(x), yi = f32_as_usize(y);
    const Col value = xyz_col(sv.texels[sv.env_texel_off + yi * sv.env_w + xi]);
    const float p = env_cdf_pdf(sv, xi, yi);
    float sp, cp, st, ct;
    spec_sincos((2.0f * RL_PI / (float)sv.env_w) * x, &sp, &cp);
    spec_sincos((RL_PI / (float)sv.env_h) * y, &st, &ct);
    *d = V3{st * cp, st * sp, ct};
    if (st == 0.0f) {
        *color = Col{0.0f, 0.0f, 0.0f};
        *pdf = 0.0f;
    } else {
        *color = value;
        *pdf = p / (2.0f * (RL_PI * RL_PI) * st);
    }
}
// ---- LightSamplerATS: importance-driven descent of the light tree (emitter.rs:1034-1108, 1319-1399) ----------------------------
#define RL_EPSILON_ATS 0.0001f
RL_HD float ats_cos_sub_clamped(float sin_a, float cos_a, float sin_b, float cos_b) { return cos_a > cos_b ? 1.0f : cos_a * cos_b + sin_a * sin_b; }
RL_HD float ats_sin_sub_clamped(float sin_a, float cos_a, float sin_b, float cos_b) { return cos_a > cos_b ? 1.0f : sin_a * cos_b - cos_a * sin_b; } // (1.0, sic: :1061-1067)
// LightBounds::importance_point(p, n) of node `id`
RL_HD float ats_importance_point(const float4 *nodes, uint32_t id, V3 p, V3 n, bool has_n) {
    const float4 c4 = nodes[4 * id], w4 = nodes[4 * id + 1], k4 = nodes[4 * id + 2];
    const V3 pc = xyz(c4), w = xyz(w4);
    const float radius = c4.w, phi = w4.w, cos_theta_o = k4.x, cos_theta_e = k4.y;
    const V3 pv = p - pc;
    const float d2 = fmaxf(dot(pv, pv), RL_EPSILON_ATS);
    const V3 wi = normalize(pv);
    float cos_theta = dot(w, wi);
    if (k4.z != 0.0f) cos_theta = fabsf(cos_theta);
    const float sin_theta = sqrtf(fmaxf(1.0f - cos_theta * cos_theta, 0.0f));
    // DirectionCone::subtended_directions(&aabb, p).cos_theta (:840-855)
    float cos_theta_u;
    if (dot(pv, pv) < radius * radius) cos_theta_u = -1.0f;
    else {
        const V3 cp = pc - p;
        const float sin_theta_max_2 = radius * radius / dot(cp, cp);
        cos_theta_u = sqrtf(fmaxf(1.0f - sin_theta_max_2, 0.0f));
    }
    const float sin_theta_u = sqrtf(fmaxf(1.0f - cos_theta_u * cos_theta_u, 0.0f));
    const float sin_theta_o = sqrtf(fmaxf(1.0f - cos_theta_o * cos_theta_o, 0.0f));
    const float cos_theta_x = ats_cos_sub_clamped(sin_theta, cos_theta, sin_theta_o, cos_theta_o);
    const float sin_theta_x = ats_sin_sub_clamped(sin_theta, cos_theta, sin_theta_o, cos_theta_o);
    const float cos_theta_p = ats_cos_sub_clamped(sin_theta_x, cos_theta_x, sin_theta_u, cos_theta_u);
    if (cos_theta_p <= cos_theta_e) return 0.0f;
    float imp = phi * cos_theta_p / d2;
    if (has_n) {
        const float cos_theta_i = fabsf(dot(wi, n));
        const float sin_theta_i = sqrtf(fmaxf(1.0f - cos_theta_i * cos_theta_i, 0.0f));
        imp *= ats_cos_sub_clamped(sin_theta_i, cos_theta_i, sin_theta_u, cos_theta_u);
    }
    return fmaxf(imp, 0.0f);
}
RL_HD float ats_prob_left(const float4 *nodes, uint32_t left, uint32_t right, V3 p, V3 n, bool has_n) {
    const float imp_left = ats_importance_point(nodes, left, p, n, has_n), imp_right = ats_importance_point(nodes, right, p, n, has_n);
    const float imp_total = imp_left + imp_right;
    return (imp_left == 0.0f && imp_right == 0.0f) ? 0.5f : imp_left / imp_total;
}
// LightSamplerATS::sample (:1361-1399): returns the global triangle index of the chosen proxy
RL_HD uint32_t ats_sample(const SceneView &sv, float r, V3 p, V3 n, float *pdf_sel_out) {
    float pdf_sel = 1.0f;
    uint32_t id = sv.ats_root;
    for (;;) {
        const float4 t = sv.ats_nodes[4 * id + 3];
        const uint32_t left = f2u(t.x), right = f2u(t.y);
        if (left == 0xffffffffu && right == 0xffffffffu) {
            *pdf_sel_out = pdf_sel;
            return f2u(t.w);
        }
        const float prob_left = ats_prob_left(sv.ats_nodes, left, right, p, n, true);
        if (r < prob_left) {
            r = r / prob_left;
            id = left;
            pdf_sel *= prob_left;
        } else {
            r = (r - prob_left) / (1.0f - prob_left);
            id = right;
            pdf_sel *= 1.0f - prob_left;
        }
    }
}
// LightSamplerATS::pdf (:1319-1359): the product of the branch probabilities from the proxy's leaf up to the root
RL_HD float ats_pdf(const SceneView &sv, uint32_t prim, V3 p, V3 n, bool has_n) {
    uint32_t id = sv.ats_leaf_of_prim[prim];
    float pdf = 1.0f;
    for (;;) {
        const uint32_t parent = f2u(sv.ats_nodes[4 * id + 3].z);
        if (parent == 0xffffffffu) return pdf;
        const float4 t = sv.ats_nodes[4 * parent + 3];
        const uint32_t left = f2u(t.x), right = f2u(t.y);
        const float prob_left = ats_prob_left(sv.ats_nodes, left, right, p, n, has_n);
        pdf *= left == id ? prob_left : 1.0f - prob_left;
        id = parent;
    }
}
struct LightSample {
    V3 p, n, d;
    Col weight;
    float pdf;
    bool valid;
    bool discrete; // PDF::Discrete: point and directional lights (no MIS against BSDF sampling)
    bool env;      // sampled on the environment: the MIS pdf of the BSDF uses the direction recomputed from p (edge.rs:37-39)
};
// EmitterSampler::sample_light -> Mesh::direct_sample -> Mesh::sample -> sample_tri
// (emitter.rs:1604-1620, 652-688; geometry.rs:340-348, 261-337; math.rs:388-394)
// `ns`: the shading normal at x (what the strategies pass as Some(&its.n_s); only the light tree uses it).  EXTRA = the kernel carries the
// rarely used branches (environment texture, light tree): KM bit 8.
template <bool EXTRA = true>
RL_HD LightSample sample_light(const SceneView &sv, V3 x, V3 ns, float r_sel, float r, float ux, float uy) {
    const bool ats = EXTRA && sv.ats_nodes != nullptr;
    uint32_t ats_prim = 0u;
    float ats_pdf_sel = 1.0f;
    if (ats) ats_prim = ats_sample(sv, r_sel, x, ns, &ats_pdf_sel); // EmitterSampler::sample_light, ATS arm (emitter.rs:1621-1638)
    // one emitter: the cdf is {0, 1} and r_sel < 1, so the search returns 0
    uint32_t id_light = (sv.n_emitters == 1u || ats) ? 0u : cdf_sample_discrete(sv.emit_cdf, sv.n_emitters + 1, r_sel);
    float pdf_sel = ats ? ats_pdf_sel : sv.emit_cdf[id_light + 1] - sv.emit_cdf[id_light];
    float4 info = ats ? make_float4(u2f(f2u(sv.shade[4 * ats_prim].w)), 0.0f, 0.0f, 0.0f) : sv.emit_info[2 * id_light];
    if (!ats && f2u(info.x) >= 0xfffffff0u) { // PointEmitter / DirectionalLight::direct_sample (emitter.rs:197-215, 115-133)
        float4 geo = sv.emit_info[2 * id_light + 1];
        Col intensity = Col{info.y, info.z, info.w};
        LightSample ls;
        Col weight;
        ls.env = false;
        if ((f2u(info.x) & 0xfu) == 2u) { // EnvironmentLight::direct_sample, emitter.rs:474-511
            V3 dd;
            Col color = intensity;
            float pdf_dir = RL_ENV_PDF;
            if (EXTRA && sv.env_w) env_sample_direction(sv, ux, uy, &dd, &color, &pdf_dir); // luminance.sample_direction(uv), texture arm
            else dd = sample_uniform_sphere(ux, uy);
            float t;
            ls.d = dd;
            ls.discrete = false;
            ls.env = true;
            if (!bsphere_intersect(xyz(geo), geo.w, x, dd, &t)) { // "Miss bSphere": dummy record with a zero weight
                ls.p = V3{0.0f, 0.0f, 0.0f};
                ls.n = V3{0.0f, 0.0f, 0.0f};
                weight = Col{0.0f, 0.0f, 0.0f};
            } else {
                ls.p = x + dd * t;
                ls.n = normalize(xyz(geo) - ls.p);
                weight = div_checked(color, pdf_dir);
            }
            ls.weight = Col{weight.r / pdf_sel, weight.g / pdf_sel, weight.b / pdf_sel};
            ls.pdf = pdf_dir * pdf_sel;
            ls.valid = ls.pdf != 0.0f;
            return ls;
        }
        if ((f2u(info.x) & 0xfu) == 0u) {
            ls.p = xyz(geo);
            V3 dd = ls.p - x;
            float dist = magnitude(dd);
            ls.d = dd / dist;
            ls.n = V3{0.0f, 0.0f, 0.0f};
            weight = div_checked(intensity, dist * dist); // intensity / dist.powi(2)
        } else {
            V3 dir = xyz(geo);
            ls.p = x - geo.w * dir; // v - bsphere.radius * direction
            ls.n = dir;
            ls.d = -dir;
            weight = intensity;
        }
        ls.weight = Col{weight.r / pdf_sel, weight.g / pdf_sel, weight.b / pdf_sel}; // res.weight /= pdf_sel (no guard)
        ls.pdf = 1.0f * pdf_sel;                                                      // PDF::Discrete(1.0) * pdf_sel
        ls.valid = ls.pdf != 0.0f;
        ls.discrete = true;
        return ls;
    }
    uint32_t mesh = f2u(info.x), first_prim = f2u(info.y), ntris = f2u(info.z), cdf_off = f2u(info.w);
    Material mat = load_material(sv.mats, mesh);
    uint32_t prim;
    if (ats) prim = ats_prim; // direct_sample_tri(p, light_info.primitive_idx, uv) (emitter.rs:608-649)
    else {
        uint32_t tri = cdf_sample_discrete(sv.area_cdf + cdf_off, ntris + 1, r);
        prim = first_prim + tri;
    }
    V3 v0 = xyz(sv.verts[3 * prim]), v1 = xyz(sv.verts[3 * prim + 1]), v2 = xyz(sv.verts[3 * prim + 2]);
    float su0 = sqrtf(ux);
    float b0 = 1.0f - su0, b1 = uy * su0;
    V3 pos = v0 * b0 + v1 * b1 + v2 * (1.0f - b0 - b1);
    // geometry.rs:272-276.  (This equals -n_geo of the shading table exactly, but reading it instead of recomputing it
    // lengthens live ranges: 176 instead of 90 bytes of spills in k_shade and 4 % more time -- measured, so recomputed.)
    V3 n_g = normalize(cross(v2 - v0, v1 - v0));
    float4 s1 = sv.shade[4 * prim + 1];
    if (f2u(s1.w) & 1u) {
        V3 n0 = xyz(s1), n1 = xyz(sv.shade[4 * prim + 2]), n2 = xyz(sv.shade[4 * prim + 3]);
        V3 n = n0 * b0 + n1 * b1 + n2 * (1.0f - b0 - b1);
        float n_l = dot(n, n);
        if (n_l == 0.0f) n = n_g;
        else if (n_l != 1.0f) n = n / sqrtf(n_l);
        if (dot(n_g, n) < 0.0f) n_g = -n_g;
    }
    float pdf_area = mat.inv_area; // 1 / cdf.total()
    if (ats) pdf_area = 1.0f / (magnitude(cross(v1 - v0, v2 - v0)) * 0.5f); // sample_tri: PDF::Area(1.0 / area_tri) (geometry.rs:328-334)
    V3 dd = pos - x;
    float dist = magnitude(dd);
    if (dist != 0.0f) dd = dd / dist;
    LightSample ls;
    ls.p = pos;
    ls.n = n_g;
    ls.d = dd;
    ls.discrete = false;
    ls.env = false;
    const float cosl = dist != 0.0f ? fmaxf(dot(n_g, -dd), 0.0f) : 0.0f;
    const float d2 = dist * dist;
    if ((dist == 0.0f || (cosl == 0.0f && d2 > 0.0f)) && pdf_sel > 0.0f) {
        // Back-facing sample: geom = 0/d2 = 0 -> pdf = 0 -> weight = 0 -> weight / pdf_sel = +0, pdf * pdf_sel = 0.  Same
        // bits as the general path below, decided without its five IEEE divisions with a zero operand, whose slow
        // path (taken by the ~20 % of lanes that sample the light from behind) cost 8 % of the kernel's instructions.
        ls.weight = Col{0.0f, 0.0f, 0.0f};
        ls.pdf = 0.0f;
        ls.valid = false;
        return ls;
    }
    float geom = dist != 0.0f ? cosl / d2 : 0.0f;
    float pdf = geom == 0.0f ? 0.0f : pdf_area / geom;
    Col le = mat.le;
    if (EXTRA && mat.emit_kind >= 2u) { // self.emit(&sampled_pos.uv): sample_tri's uv (geometry.rs:316-325), the interpolation NORMALIZED as a 2-vector (sic)
        const float2 q0 = sv.uvs[3 * prim], q1 = sv.uvs[3 * prim + 1], q2 = sv.uvs[3 * prim + 2];
        const float b2 = 1.0f - b0 - b1;
        const float qx = q0.x * b0 + q1.x * b1 + q2.x * b2, qy = q0.y * b0 + q1.y * b1 + q2.y * b2;
        const float il = 1.0f / sqrtf(qx * qx + qy * qy);
        le = mesh_emit(sv, mat, qx * il, qy * il);
    }
    Col weight = pdf == 0.0f ? Col{0.0f, 0.0f, 0.0f} : div_checked(mul_checked(le, geom), pdf_area);
    ls.weight = pdf_sel == 1.0f ? weight : Col{weight.r / pdf_sel, weight.g / pdf_sel, weight.b / pdf_sel}; // x / 1 == x
    ls.pdf = pdf * pdf_sel;
    ls.valid = ls.pdf != 0.0f;
    return ls;
}
// EmitterSampler::direct_pdf (emitter.rs:1566-1575) -> Mesh::direct_pdf (emitter.rs:571-579)
// With the light tree (emitter.rs:1573-1601): Mesh::direct_pdf_tri (:581-589, the TRIANGLE's area pdf) times LightSamplerATS::pdf evaluated
// from `o` with the normal the caller has (`path`: None, emitters.rs:52-57; `direct`: Some(&its.n_s), direct.rs:158-165).
template <bool EXTRA = true>
RL_HD float direct_pdf(const SceneView &sv, const Material &light_mat, uint32_t prim, V3 o, V3 p, V3 n, V3 dir, V3 ns, bool has_ns) {
    const bool ats = EXTRA && sv.ats_nodes != nullptr;
    float cos_light = fmaxf(dot(n, -dir), 0.0f);
    float v;
    if (cos_light == 0.0f) v = 0.0f;
    else {
        V3 po = p - o;
        float geom = cos_light / dot(po, po);
        float pdf_area = light_mat.inv_area;
        if (ats) { // Mesh::pdf_tri, geometry.rs:226-234
            const V3 v0 = xyz(sv.verts[3 * prim]), v1 = xyz(sv.verts[3 * prim + 1]), v2 = xyz(sv.verts[3 * prim + 2]);
            pdf_area = 1.0f / (magnitude(cross(v1 - v0, v2 - v0)) * 0.5f);
        }
        v = pdf_area / geom;
    }
    return v * (ats ? ats_pdf(sv, prim, o, ns, has_ns) : light_mat.pdf_sel);
}

// mis_weight, integrators/mod.rs:462-478 (power heuristic, `direct` only)
RL_HD float mis_weight_power(float pdf_a, float pdf_b) {
    if (pdf_a == 0.0f) return 0.0f;
    if (!finite_f(pdf_a) || !finite_f(pdf_b)) return 0.0f;
    float w = (pdf_a * pdf_a) / ((pdf_a * pdf_a) + (pdf_b * pdf_b));
    return finite_f(w) ? w : 0.0f;
}

// ---- integrator parameters ------------------------------------------------------------------------
struct IntegParams {
    uint32_t kind;     // 0 path, 1 direct, 2 ao
    float ao_max_distance;         // < 0: None
    uint32_t ao_normal_correction;
    int32_t min_depth, max_depth, rr_depth;
    uint32_t strategy; // 0 all, 1 bsdf, 2 emitter
    uint32_t single_scattering;
    uint32_t nb_bsdf_samples, nb_light_samples;
    uint64_t seed_h;   // seed_hash(seed)
    uint32_t sample_base; // first sample index of this batch
    uint32_t npix;        // pixels owned by this rank
    uint32_t npix_mul, npix_shift; // id / npix == umulhi(id, npix_mul) >> npix_shift for every path id of the batch (0: plain division)
    uint32_t img_w;
};

// ---- one wavefront step of the `path` integrator for one path ------------------------------------
// Input: the ray that was traced (o, d), its hit, the path state.  Output: radiance to add
// now (arrival emission), an optional next ray + state, an optional shadow segment + the
// radiance it carries.  Mirrors generate()/evaluate() of the reference in streaming form
// (identical to oracle.cpp path_compute_pixel_stream).
struct PathState {
    Col T;            // throughput entering the vertex that this ray hits
    float pdf_prev;   // solid-angle pdf of the direction that produced this ray
    uint32_t path_id; // s_local * npix + local pixel
    uint32_t depth;   // generate() depth at which this ray was sampled (1 = sensor)
    uint32_t rng_n;   // random numbers consumed so far
};
struct StepOut {
    Col add;          // arrival emission (already throughput- and MIS-weighted)
    bool has_add;
    bool alive;       // next ray valid
    V3 next_o, next_d;
    PathState next;
    bool nee_sampled; // the reference would call Acceleration::visible here
    bool shadow;      // a shadow segment must be traced
    V3 sh_p0, sh_p1;
    Col sh_contrib;
};
RL_HD uint32_t umulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}
// path id = s_local * npix + local pixel
RL_HD void ip_split(const IntegParams &ip, uint32_t id, uint32_t *s_local, uint32_t *lp) {
    const uint32_t q = ip.npix_mul ? (umulhi32(id, ip.npix_mul) >> ip.npix_shift) : id / ip.npix;
    *s_local = q;
    *lp = id - q * ip.npix;
}
RL_HD bool ip_expand(const IntegParams &ip, uint32_t depth) { return ip.max_depth < 0 ? true : depth < (uint32_t)ip.max_depth; }
RL_HD bool ip_add_ok(const IntegParams &ip, uint32_t curr_depth) { return ip.min_depth < 0 ? true : curr_depth >= (uint32_t)ip.min_depth; }

template <uint32_t KM = RL_KM_ALL>
RL_HD void path_step(const SceneView &sv, const IntegParams &ip, V3 o, V3 d, const HitRec &hit, const PathState &st, uint32_t pixel,
                     uint32_t sample, StepOut *out) {
    out->has_add = false;
    out->alive = false;
    out->shadow = false;
    out->nee_sampled = false;
    out->add = Col{0.0f, 0.0f, 0.0f};
    if (hit.prim == RL_MISS) {
        // Edge without a next vertex (edge.rs:93-127): contributes weight * rr * environment luminance (edge.rs:208), with the
        // same gates and MIS as an emitter hit (path.rs:37-111, 152-165); the light strategy's pdf is pdf_emitter's
        // environment arm (emitters.rs:18-46): constant 1 / 4 pi times the selection probability.
        if (!sv.env_on) return;
        const bool envtex = RL_HAS(KM, 8) && sv.env_w != 0u; // scene.enviroment_luminance(d) = env.eval(d): constant, or the texel d points at
        const Col env_l = envtex ? env_eval(sv, d) : sv.env_color;
        if (st.depth == 1u) {
            if (ip_add_ok(ip, 0u) && !is_zero(env_l)) {
                out->add = env_l;
                out->has_add = true;
            }
        } else if (ip.single_scattering == 0u && ip_add_ok(ip, st.depth - 1u) && ip.strategy != 2u) {
            Col contrib = st.T * env_l;
            if (!is_zero(contrib)) {
                float w = 1.0f;
                if (ip.strategy == 0u && !(f2u(st.pdf_prev) >> 31)) {
                    float pl = (envtex ? env_pdf(sv, d) : RL_ENV_PDF) * sv.env_pdf_sel;
                    w = st.pdf_prev / (st.pdf_prev + pl);
                }
                out->add = mul_checked(contrib, w);
                out->has_add = true;
            }
        }
        return;
    }
    float4 s0 = sv.shade[4 * hit.prim], s1 = sv.shade[4 * hit.prim + 1], s2 = sv.shade[4 * hit.prim + 2], s3 = sv.shade[4 * hit.prim + 3];
    uint32_t mesh = f2u(s0.w);
    Material mat = load_material(sv.mats, mesh);
    if (RL_HAS(KM, 8) && (sv.tex || sv.emit_var)) apply_textures(sv, mat, hit.prim, f2u(s1.w), hit.u, hit.v);
    Surface its = fill_intersection<KM>(sv, mat, hit.prim, mesh, s0, s1, s2, s3, hit.t, hit.u, hit.v, o, d);
    const bool mute = ip.single_scattering != 0u;
    const bool smooth = mat_is_smooth<KM>(mat); // no light sampling at this vertex (emitters.rs:110-112), no draws either
    const bool use_nee = (ip.strategy == 0u || ip.strategy == 2u) && !smooth;
    // ---- emission carried by the arriving edge --------------------------------------------------
    if (st.depth == 1u) { // sensor edge: un-weighted (path.rs:152-165)
        if (ip_add_ok(ip, 0u) && dot(its.n_s, -d) >= 0.0f && mat.is_light && !is_zero(mat.le)) {
            out->add = mat.le;
            out->has_add = true;
        }
    } else if (!mute && ip_add_ok(ip, st.depth - 1u) && ip.strategy != 2u) {
        if (dot(its.n_s, -d) >= 0.0f && mat.is_light) {
            Col contrib = st.T * mat.le;
            if (!is_zero(contrib)) {
                float w = 1.0f;
                // balance heuristic against light sampling (path.rs:78-99); a negative pdf_prev marks an edge without
                // MIS: PDF::Discrete (delta lobe), or sampled at a smooth vertex where the light strategy has no pdf
                if (ip.strategy == 0u && !(f2u(st.pdf_prev) >> 31)) {
                    float pl = direct_pdf<RL_HAS(KM, 8) != 0u>(sv, mat, hit.prim, o, its.p, its.n_g, d, V3{0.0f, 0.0f, 0.0f}, false);
                    w = st.pdf_prev / (st.pdf_prev + pl);
                }
                out->add = mul_checked(contrib, w);
                out->has_add = true;
            }
        }
    }
    // ---- expand this vertex ----------------------------------------------------------------------
    uint32_t depth = st.depth + 1u;
    if (!ip_expand(ip, depth)) return;
#if RL_PAIR_SAMPLER
    PairSampler smp = make_pair_sampler(ip.seed_h, pixel, sample, st.rng_n);
#else
    Sampler smp = make_sampler(ip.seed_h, pixel, sample, st.rng_n);
#endif
    // directional strategy: BSDF sample + Russian roulette (directional.rs:44-107)
    {
#if RL_PAIR_SAMPLER
        float sx = smp.even();
        float sy = smp.odd();
#else
        float sx = smp.next();
        float sy = smp.next();
#endif
        Col bw;
        V3 wo;
        float bpdf;
        bool discrete;
        if (bsdf_sample<KM>(mat, its.wi, sx, sy, &bw, &wo, &bpdf, &discrete)) {
            V3 d_out = to_world(its.frame, wo);
            Col Tn = st.T * bw;
            if (!is_zero(Tn)) {
                bool do_rr = ip.rr_depth < 0 ? true : (uint32_t)ip.rr_depth <= depth;
                bool survive = true;
                float rr_weight = 1.0f;
                if (do_rr) {
                    float q = fminf(channel_max(Tn), 0.95f);
#if RL_PAIR_SAMPLER
                    if (q < smp.even()) survive = false;
#else
                    if (q < smp.next()) survive = false;
#endif
                    else rr_weight = 1.0f / q;
                }
                if (survive) {
                    Tn.r *= rr_weight;
                    Tn.g *= rr_weight;
                    Tn.b *= rr_weight;
                    out->alive = true;
                    out->next_o = its.p;
                    out->next_d = d_out;
                    out->next.T = Tn;
                    out->next.pdf_prev = (discrete || smooth) ? u2f(f2u(bpdf) | 0x80000000u) : bpdf;
                    out->next.path_id = st.path_id;
                    out->next.depth = depth;
                }
            }
        }
    }
    // light sampling strategy (emitters.rs:108-175); runs even when the bounce died
    if (use_nee) {
#if RL_PAIR_SAMPLER
        smp.begin(smp.n); // the roulette draw above is conditional: re-begin at the count reached
        float r_sel = smp.even();
        float r = smp.odd();
        float ux = smp.even();
        float uy = smp.odd();
#else
        float r_sel = smp.next();
        float r = smp.next();
        float ux = smp.next();
        float uy = smp.next();
#endif
        out->nee_sampled = true;
        LightSample ls = sample_light<RL_HAS(KM, 8) != 0u>(sv, its.p, its.n_s, r_sel, r, ux, uy);
        if (ls.valid && !mute && ip_add_ok(ip, st.depth) && ip.strategy != 1u) {
            V3 wo = to_local(its.frame, ls.d);
            Col f;
            float pb;
            bsdf_eval_pdf<KM>(mat, its.wi, wo, &f, &pb);
            if (ls.env) { // the edge direction is recomputed from the two positions (Edge::from_vertex, edge.rs:37-39); for mesh lights
                          // that IS ls.d, for the environment (p = x + d t) it differs in the last bits
                V3 de = ls.p - its.p;
                float dist = magnitude(de);
                de = de / dist;
                pb = bsdf_pdf<KM>(mat, its.wi, to_local(its.frame, de));
            }
            Col contrib = st.T * (ls.weight * f);
            if (!is_zero(contrib)) {
                float w = 1.0f;
                if (ip.strategy == 0u && !ls.discrete) w = ls.pdf / (pb + ls.pdf); // a PDF::Discrete light edge has no MIS (path.rs:80)
                out->shadow = true;
                out->sh_p0 = its.p;
                out->sh_p1 = ls.p;
                out->sh_contrib = mul_checked(contrib, w);
            }
        }
    }
    out->next.rng_n = smp.n;
}

// ---- `direct` integrator (IntegratorDirect::compute_pixel, direct.rs:21-233) ------------------------
// Stage 1 runs on the primary hit: emission, then nb_light_samples shadow segments and
// nb_bsdf_samples extension rays.  Stage 2 runs on the hit of an extension ray.  Every
// contribution of one pixel sample goes to its own radiance slot (0 = emission, 1..nl = light
// samples, nl+1.. = BSDF samples) and the slots are summed in that order, which is the order of
// the reference's `l_i +=` statements.
struct DirectCtx {
    Surface its;
    Material mat;
    Sampler smp;
    float wb, wl; // weight_nb_bsdf, weight_nb_light (direct.rs:48-57)
    bool ok;      // primary ray hit a surface seen from its front side
    bool env_primary; // the primary ray escaped into the environment: the sample's value is `emit`
    Col emit;
};
template <uint32_t KM = RL_KM_ALL>
RL_HD void direct_begin(const SceneView &sv, const IntegParams &ip, V3 o, V3 d, const HitRec &hit, uint32_t rng_n, uint32_t pixel, uint32_t sample,
                        DirectCtx *cx) {
    cx->ok = false;
    cx->emit = Col{0.0f, 0.0f, 0.0f};
    cx->env_primary = false;
    if (hit.prim == RL_MISS) { // return scene.enviroment_luminance(ray.d) (direct.rs:33-36)
        if (sv.env_on) cx->emit = (RL_HAS(KM, 8) && sv.env_w) ? env_eval(sv, d) : sv.env_color, cx->env_primary = true;
        return;
    }
    float4 s0 = sv.shade[4 * hit.prim], s1 = sv.shade[4 * hit.prim + 1], s2 = sv.shade[4 * hit.prim + 2], s3 = sv.shade[4 * hit.prim + 3];
    uint32_t mesh = f2u(s0.w);
    cx->mat = load_material(sv.mats, mesh);
    if (RL_HAS(KM, 8) && (sv.tex || sv.emit_var)) apply_textures(sv, cx->mat, hit.prim, f2u(s1.w), hit.u, hit.v);
    cx->its = fill_intersection<KM>(sv, cx->mat, hit.prim, mesh, s0, s1, s2, s3, hit.t, hit.u, hit.v, o, d);
    if (cx->its.wi.z <= 0.0f) return; // its.cos_theta() <= 0 (direct.rs:40-42)
    cx->ok = true;
    if (cx->mat.is_light) cx->emit = cx->mat.le;
    cx->wb = ip.nb_bsdf_samples == 0u ? 0.0f : 1.0f / (float)ip.nb_bsdf_samples;
    cx->wl = ip.nb_light_samples == 0u ? 0.0f : 1.0f / (float)ip.nb_light_samples;
    cx->smp = make_sampler(ip.seed_h, pixel, sample, rng_n);
}
// One light sample (direct.rs:63-129).  Returns true when a shadow segment must be traced;
// *valid tells whether the reference would have called Acceleration::visible.
template <uint32_t KM = RL_KM_ALL>
RL_HD bool direct_light_sample(const SceneView &sv, DirectCtx *cx, V3 *p1, Col *contrib, bool *valid) {
    float r_sel = cx->smp.next();
    float r = cx->smp.next();
    float ux = cx->smp.next();
    float uy = cx->smp.next();
    LightSample ls = sample_light<RL_HAS(KM, 8) != 0u>(sv, cx->its.p, cx->its.n_s, r_sel, r, ux, uy);
    *valid = ls.valid;
    if (!ls.valid) return false;
    if (mat_is_smooth<KM>(cx->mat)) return false; // direct.rs:74-77: visible() is still called (valid stays true), nothing is added
    V3 wo = to_local(cx->its.frame, ls.d);
    float pdf_bsdf;
    Col f;
    bsdf_eval_pdf<KM>(cx->mat, cx->its.wi, wo, &f, &pdf_bsdf);
    float weight_light = ls.discrete ? 1.0f : mis_weight_power(ls.pdf * cx->wl, pdf_bsdf * cx->wb); // direct.rs:106-110
    Col c = mul_checked(mul_plain(weight_light, f), cx->wl) * ls.weight;
    *p1 = ls.p;
    *contrib = c;
    return !is_zero(c);
}
// ---- IntegratorAO::compute_pixel (ao.rs:20-72) on the same two stages: stage 1 = primary hit -> one cosine-distributed
// extension ray (flipped when normal_correction and the surface is seen from behind), stage 2 = 1 when the ray escapes
// or, with a max_distance, when the next surface is further away.
RL_HD void ao_begin(const SceneView &sv, const IntegParams &ip, V3 o, V3 d, const HitRec &hit, uint32_t rng_n, uint32_t pixel, uint32_t sample, DirectCtx *cx) {
    cx->ok = false;
    cx->env_primary = false;
    cx->emit = Col{0.0f, 0.0f, 0.0f};
    if (hit.prim == RL_MISS) return;
    float4 s0 = sv.shade[4 * hit.prim], s1 = sv.shade[4 * hit.prim + 1], s2 = sv.shade[4 * hit.prim + 2], s3 = sv.shade[4 * hit.prim + 3];
    uint32_t mesh = f2u(s0.w);
    cx->mat = load_material(sv.mats, mesh);
    cx->its = fill_intersection(sv, cx->mat, hit.prim, mesh, s0, s1, s2, s3, hit.t, hit.u, hit.v, o, d);
    if (ip.ao_normal_correction == 0u && cx->its.wi.z <= 0.0f) return;
    cx->ok = true;
    cx->smp = make_sampler(ip.seed_h, pixel, sample, rng_n);
}
RL_HD bool ao_sample(const IntegParams &ip, DirectCtx *cx, V3 *dir) {
    const bool flipped = ip.ao_normal_correction != 0u && cx->its.wi.z <= 0.0f;
    float sx = cx->smp.next();
    float sy = cx->smp.next();
    V3 d_local = cosine_sample_hemisphere(sx, sy);
    *dir = to_world(cx->its.frame, flipped ? -d_local : d_local);
    return true;
}
RL_HD bool ao_finish(const IntegParams &ip, const HitRec &hit, Col *contrib) {
    const bool open = hit.prim == RL_MISS || (ip.ao_max_distance >= 0.0f && hit.t > ip.ao_max_distance);
    *contrib = Col{1.0f, 1.0f, 1.0f};
    return open;
}
// One BSDF sample (direct.rs:135-144): the extension ray and what stage 2 needs (weight, pdf).
template <uint32_t KM = RL_KM_ALL>
RL_HD bool direct_bsdf_sample(DirectCtx *cx, V3 *dir, Col *weight, float *pdf) {
    float sx = cx->smp.next();
    float sy = cx->smp.next();
    V3 wo;
    bool discrete;
    if (!bsdf_sample<KM>(cx->mat, cx->its.wi, sx, sy, weight, &wo, pdf, &discrete)) return false;
    if (discrete) *pdf = u2f(f2u(*pdf) | 0x80000000u); // PDF::Discrete -> weight_bsdf = 1 in stage 2 (direct.rs:170)
    *dir = to_world(cx->its.frame, wo);
    return true;
}
// Stage 2 (direct.rs:145-181): the extension ray hit something; contribution if it is a light.
// (ns, has_ns: the first vertex' shading normal, which the light tree's pdf of the second hit takes: Some(&its.n_s), direct.rs:158-165)
RL_HD bool direct_finish(const SceneView &sv, const IntegParams &ip, V3 o, V3 d, const HitRec &hit, Col bsdf_weight, float bsdf_pdf_v, Col *contrib,
                         V3 ns = V3{0.0f, 0.0f, 0.0f}, bool has_ns = false) {
    if (hit.prim == RL_MISS) { // direct.rs:183-227: the BSDF sample escaped: MIS against sampling the environment
        if (!sv.env_on) return false;
        float wb = ip.nb_bsdf_samples == 0u ? 0.0f : 1.0f / (float)ip.nb_bsdf_samples;
        float wl = ip.nb_light_samples == 0u ? 0.0f : 1.0f / (float)ip.nb_light_samples;
        float weight_bsdf = 1.0f;
        if (!(f2u(bsdf_pdf_v) >> 31)) weight_bsdf = mis_weight_power(bsdf_pdf_v * wb, ((sv.env_w ? env_pdf(sv, d) : RL_ENV_PDF) * sv.env_pdf_sel) * wl);
        *contrib = mul_checked(mul_plain(weight_bsdf, bsdf_weight) * (sv.env_w ? env_eval(sv, d) : sv.env_color), wb);
        return true;
    }
    float4 s0 = sv.shade[4 * hit.prim], s1 = sv.shade[4 * hit.prim + 1], s2 = sv.shade[4 * hit.prim + 2], s3 = sv.shade[4 * hit.prim + 3];
    uint32_t mesh = f2u(s0.w);
    Material mat = load_material(sv.mats, mesh);
    if (!mat.is_light) return false;
    if (mat.emit_kind >= 2u) apply_textures(sv, mat, hit.prim, f2u(s1.w), hit.u, hit.v); // next_its.mesh.emit(&next_its.uv) (direct.rs:179)
    Surface nx = fill_intersection(sv, mat, hit.prim, mesh, s0, s1, s2, s3, hit.t, hit.u, hit.v, o, d);
    if (!(dot(nx.n_g, -d) > 0.0f)) return false;
    float wb = ip.nb_bsdf_samples == 0u ? 0.0f : 1.0f / (float)ip.nb_bsdf_samples;
    float wl = ip.nb_light_samples == 0u ? 0.0f : 1.0f / (float)ip.nb_light_samples;
    float weight_bsdf = 1.0f;
    if (!(f2u(bsdf_pdf_v) >> 31)) {
        float light_pdf = direct_pdf(sv, mat, hit.prim, o, nx.p, nx.n_g, d, ns, has_ns);
        weight_bsdf = mis_weight_power(bsdf_pdf_v * wb, light_pdf * wl);
    }
    *contrib = mul_checked(mul_plain(weight_bsdf, bsdf_weight) * mat.le, wb);
    return true;
}

} // namespace rl
