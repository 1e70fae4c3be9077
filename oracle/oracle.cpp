// oracle.cpp -- CPU restatement of rustlight@864df34's `path` / `direct` hot path.
//
// TEST INFRASTRUCTURE, NOT PRODUCT.  PARITY UNPINNED (see oracle.h).
//
// Every section names the reference file:line it follows.  The layout deliberately mirrors
// the reference (AoS records, recursive BVH, explicit path graph, trait-like virtual calls),
// which is the opposite of the GPU implementation, so that agreement between the two is
// evidence and not tautology.  Build with -O2 -ffp-contract=off -fno-fast-math: every f32
// operation below is a single IEEE-754 operation in the order written.
//
//   structure.rs  : PDF 20-58,71-94 | Color 106-380 | Ray 697-732 | AABB 760-878 | Intersection 924-1059
//   math.rs       : sampling 37-72 | Frame 357-384 | uniform_sample_triangle 388-394 | Distribution1D 398-487
//   geometry.rs   : Mesh::new 122-182 | emit 184-206 | pdf 223-225 | sample_tri/sample 261-348
//                   intersection_tri 358-410 | compute_aabb_tri 423-439
//   accel.rs      : NaiveAcceleration 22-77 | BVHAccel build 107-240, traversal 243-288, trace/visible 292-343
//   bsdfs/        : diffuse.rs 10-87 | phong.rs 14-136 | mod.rs 124-161
//   emitter.rs    : LightSampling 10-44 | impl Emitter for Mesh 570-688 | EmitterSampler 1491-1647
//   scene.rs      : build_emitters 53-123
//   camera.rs     : Camera::new 31-67 | generate 81-91
//   samplers/     : mod.rs 3-9 | independent.rs 5-34  (+ rand 0.8.5 SmallRng, restated from its published algorithm)
//   paths/        : path.rs 56-73 | vertex.rs 61-82 | edge.rs 27-210 | strategies/{mod,directional,emitters}.rs
//   integrators/  : mod.rs 351-478 | explicit/path.rs 27-238 | direct.rs 20-233
#include "oracle.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstring>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace {

constexpr float EPSILON = 0.0001f;                     // lib.rs:51
constexpr float F32_MAX = std::numeric_limits<float>::max();
constexpr float PI = 3.14159265358979323846264338327950288f;
constexpr float FRAC_PI_2 = 1.57079632679489661923132169163975144f;
constexpr float FRAC_PI_4 = 0.785398163397448309615660845819875721f;
constexpr float FRAC_1_PI = 0.318309886183790671537767526745028724f;

// ============================================================================================
// cgmath 0.18 vector arithmetic (crate not vendored; restated: SURVEY.md §8c)
// ============================================================================================
struct V3 {
    float x, y, z;
};
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator*(float s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline V3 operator/(V3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; } // (xx + yy) + zz
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline float magnitude2(V3 a) { return dot(a, a); }
inline float magnitude(V3 a) { return std::sqrt(dot(a, a)); }
inline V3 normalize(V3 a) { return a * (1.0f / magnitude(a)); } // normalize_to(1): v * (1/|v|)
inline float comp(V3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
inline V3 load3(const float *p) { return {p[0], p[1], p[2]}; }
inline void store3(float *p, V3 v) { p[0] = v.x, p[1] = v.y, p[2] = v.z; }
// Rust f32::max / f32::min: if one operand is NaN the other is returned == fmaxf / fminf
inline float rmax(float a, float b) { return std::fmax(a, b); }
inline float rmin(float a, float b) { return std::fmin(a, b); }

struct M4 { // column-major, m[4*c+r]
    float m[16];
    float at(int c, int r) const { return m[4 * c + r]; }
};
// Matrix4 * Vector4 = c0*x + c1*y + c2*z + c3*w
inline void m4_mul_v4(const M4 &m, const float v[4], float out[4]) {
    for (int r = 0; r < 4; r++) out[r] = ((m.at(0, r) * v[0] + m.at(1, r) * v[1]) + m.at(2, r) * v[2]) + m.at(3, r) * v[3];
}
inline V3 transform_point(const M4 &m, V3 p) { // Point3::from_homogeneous(m * p.to_homogeneous())
    float v[4] = {p.x, p.y, p.z, 1.0f}, h[4];
    m4_mul_v4(m, v, h);
    float iw = 1.0f / h[3];
    return {h[0] * iw, h[1] * iw, h[2] * iw};
}
inline V3 transform_vector(const M4 &m, V3 d) { // (m * d.extend(0)).truncate()
    float v[4] = {d.x, d.y, d.z, 0.0f}, h[4];
    m4_mul_v4(m, v, h);
    return {h[0], h[1], h[2]};
}

// ============================================================================================
// "spec" transcendental functions: f64 polynomials, identical op sequence on
// the GPU.  Used when math_mode == ORC_MATH_SPEC; ORC_MATH_LIBM calls glibc like Rust does.
// ============================================================================================
inline void spec_sincos(float x, float *s, float *c) {
    // f32: three-piece Cody-Waite reduction by pi/2, Cephes sinf / cosf kernels; one fmaf / mul per step
    const float TWO_OVER_PI = 0.636619772f;
    const float PIO2_A = 1.5703125f, PIO2_B = 4.837512969970703125e-4f, PIO2_C = 7.54978995489188e-8f;
    float fn = std::rint(x * TWO_OVER_PI);
    int n = (int)fn;
    float y = std::fma(fn, -PIO2_A, x);
    y = std::fma(fn, -PIO2_B, y);
    y = std::fma(fn, -PIO2_C, y);
    float y2 = y * y;
    float ps = -1.9515295891e-4f;
    ps = std::fma(ps, y2, 8.3321608736e-3f);
    ps = std::fma(ps, y2, -1.6666654611e-1f);
    float sy = std::fma(y * y2, ps, y);
    float pc = 2.443315711809948e-5f;
    pc = std::fma(pc, y2, -1.388731625493765e-3f);
    pc = std::fma(pc, y2, 4.166664568298827e-2f);
    float cy = std::fma(y2 * y2, pc, std::fma(-0.5f, y2, 1.0f));
    switch (n & 3) {
    case 0: *s = sy, *c = cy; break;
    case 1: *s = cy, *c = -sy; break;
    case 2: *s = -sy, *c = -cy; break;
    default: *s = -cy, *c = sy; break;
    }
}
inline double spec_log2(double x) { // x > 0, normal
    uint64_t bits;
    std::memcpy(&bits, &x, 8);
    int e = (int)((bits >> 52) & 0x7ff) - 1023;
    uint64_t mb = (bits & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL;
    double m;
    std::memcpy(&m, &mb, 8);
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
inline double spec_exp2(double t) {
    if (t < -1000.0) return 0.0;
    if (t > 1000.0) return std::numeric_limits<double>::infinity();
    double k = std::floor(t + 0.5);
    double r = (t - k) * 0.69314718055994530942; // |r| <= 0.3466
    double p = 1.0 / 6227020800.0;               // 1/13!
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
    // scale by 2^ki in two steps so that subnormal results round once
    int k1 = ki / 2, k2 = ki - k1;
    uint64_t b1 = (uint64_t)(k1 + 1023) << 52, b2 = (uint64_t)(k2 + 1023) << 52;
    double s1, s2;
    std::memcpy(&s1, &b1, 8);
    std::memcpy(&s2, &b2, 8);
    return (p * s1) * s2;
}
inline float spec_powf(float x, float y) { // x >= 0
    if (y == 0.0f) return 1.0f;
    if (x == 0.0f) return y > 0.0f ? 0.0f : std::numeric_limits<float>::infinity();
    if (x == 1.0f) return 1.0f;
    if (!(x > 0.0f)) return std::numeric_limits<float>::quiet_NaN();
    return (float)spec_exp2((double)y * spec_log2((double)x));
}
// atan2 / acos of the SPEC math mode: Cephes atanf / asinf kernels in f32, one fmaf / mul / div / sqrt per step;
// typed independently of rl_device.cuh, same operation sequence.
inline float spec_atan_nonneg(float x) {
    float base, z;
    if (x > 2.414213562373095f) base = 1.5707963267948966f, z = -(1.0f / x);
    else if (x > 0.4142135623730950f) base = 0.7853981633974483f, z = (x - 1.0f) / (x + 1.0f);
    else base = 0.0f, z = x;
    const float zz = z * z;
    float poly = std::fmaf(8.05374449538e-2f, zz, -1.38776856032e-1f);
    poly = std::fmaf(poly, zz, 1.99777106478e-1f);
    poly = std::fmaf(poly, zz, -3.33329491539e-1f);
    return base + std::fmaf(poly * zz, z, z);
}
inline float spec_atan2(float y, float x) {
    if (x != x || y != y) return x + y;
    if (x == 0.0f) return y > 0.0f ? 1.5707963267948966f : (y < 0.0f ? -1.5707963267948966f : 0.0f);
    const float q = y / x;
    const float a = spec_atan_nonneg(std::fabs(q));
    const float at = q < 0.0f ? -a : a;
    if (x > 0.0f) return at;
    return std::signbit(y) ? at - 3.14159265358979323846f : at + 3.14159265358979323846f; // atan2(-0, x < 0) = -pi
}
inline float spec_asin_nonneg(float a) {
    if (a < 1e-4f) return a;
    const bool big = a > 0.5f;
    float z, x;
    if (big) z = 0.5f * (1.0f - a), x = std::sqrt(z);
    else x = a, z = x * x;
    float poly = std::fmaf(4.2163199048e-2f, z, 2.4181311049e-2f);
    poly = std::fmaf(poly, z, 4.5470025998e-2f);
    poly = std::fmaf(poly, z, 7.4953002686e-2f);
    poly = std::fmaf(poly, z, 1.6666752422e-1f);
    float r = std::fmaf(poly * z, x, x);
    if (big) r = 1.5707963267948966f - (r + r);
    return r;
}
inline float spec_acos(float x) {
    if (x != x) return x;
    if (x < -0.5f) return 3.14159265358979323846f - 2.0f * spec_asin_nonneg(std::sqrt(0.5f * (1.0f + x)));
    if (x > 0.5f) return 2.0f * spec_asin_nonneg(std::sqrt(0.5f * (1.0f - x)));
    const float a = spec_asin_nonneg(std::fabs(x));
    return 1.5707963267948966f - (x < 0.0f ? -a : a);
}
struct Math {
    uint32_t mode;
    float atan2(float y, float x) const { return mode == ORC_MATH_LIBM ? std::atan2(y, x) : spec_atan2(y, x); }
    float acos(float x) const { return mode == ORC_MATH_LIBM ? std::acos(x) : spec_acos(x); }
    float sin(float x) const {
        float sn, cs;
        sincos(x, &sn, &cs);
        return sn;
    }
    void sincos(float x, float *s, float *c) const {
        if (mode == ORC_MATH_LIBM) {
            *s = std::sin(x);
            *c = std::cos(x);
        } else spec_sincos(x, s, c);
    }
    float powf(float x, float y) const { return mode == ORC_MATH_LIBM ? std::pow(x, y) : spec_powf(x, y); }
    float exp(float x) const { return mode == ORC_MATH_LIBM ? std::exp(x) : (float)spec_exp2((double)x * 1.4426950408889634074); }
    float ln(float x) const {
        if (mode == ORC_MATH_LIBM) return std::log(x);
        if (x == 0.0f) return -std::numeric_limits<float>::infinity();
        if (!(x > 0.0f)) return std::numeric_limits<float>::quiet_NaN();
        if (!std::isfinite(x)) return x;
        return (float)(spec_log2((double)x) * 0.69314718055994530942);
    }
};

// ============================================================================================
// structure.rs: PDF, Color
// ============================================================================================
struct PDF { // structure.rs:19-24
    enum Kind { SolidAngle, Area, Discrete } kind;
    float v;
    bool is_zero() const { return v == 0.0f; }                  // :72-76
    float value() const { return v; }                           // :78-82
    PDF operator*(float o) const { return PDF{kind, v * o}; }   // :85-94
    PDF as_solid_angle_geom(float g_ad) const {                 // :27-39
        if (kind == SolidAngle) return *this;
        if (g_ad == 0.0f) return PDF{SolidAngle, 0.0f};
        return PDF{SolidAngle, v / g_ad};
    }
};
struct Color { // structure.rs:105-110
    float r, g, b;
    static Color zero() { return {0, 0, 0}; }
    static Color one() { return {1, 1, 1}; }
    bool is_zero() const { return r == 0.0f && g == 0.0f && b == 0.0f; }     // :153-155
    float channel_max() const { return rmax(r, rmax(g, b)); }                // :169-171
    void scale(float v) { r *= v, g *= v, b *= v; }                          // :185-191
};
inline Color operator*(Color a, float o) { // :278-292 (zero if the scalar is not finite)
    if (std::isfinite(o)) return {a.r * o, a.g * o, a.b * o};
    return Color::zero();
}
inline Color operator*(float s, Color o) { return {o.r * s, o.g * s, o.b * s}; } // :294-303
inline Color operator*(Color a, Color b) { return {a.r * b.r, a.g * b.g, a.b * b.b}; } // :338-347
inline Color operator+(Color a, Color b) { return {a.r + b.r, a.g + b.g, a.b + b.b}; } // :360-369
inline Color operator/(Color a, float o) { // :249-265
    if (o == 0.0f || !std::isfinite(o)) return Color::zero();
    return {a.r / o, a.g / o, a.b / o};
}
inline void div_assign(Color &a, float o) { a.r /= o, a.g /= o, a.b /= o; } // :201-207 (no guard)
inline Color operator-(Color a, Color b) { return {a.r - b.r, a.g - b.g, a.b - b.b}; } // :349-358
inline Color operator/(Color a, Color b) { return {a.r / b.r, a.g / b.g, a.b / b.b}; } // :266-275 (no guard)
inline Color color_value(float v) { return {v, v, v}; }                                 // :122-124
inline Color safe_sqrt(Color c) { return {std::sqrt(rmax(c.r, 0.0f)), std::sqrt(rmax(c.g, 0.0f)), std::sqrt(rmax(c.b, 0.0f))}; } // :132-138

// ============================================================================================
// structure.rs: Ray, AABB
// ============================================================================================
struct Ray { // :697-702
    V3 o, d;
    float tnear, tfar;
};
inline Ray ray_new(V3 o, V3 d) { return Ray{o, d, EPSILON, F32_MAX}; } // :705-715 (assert on |d| omitted)

struct AABB { // :760-763
    V3 p_min{F32_MAX, F32_MAX, F32_MAX}, p_max{-F32_MAX, -F32_MAX, -F32_MAX}; // Default :765-772 (f32::MIN == -MAX)
    AABB union_aabb(const AABB &b) const { // :779-784
        AABB r;
        r.p_min = {rmin(p_min.x, b.p_min.x), rmin(p_min.y, b.p_min.y), rmin(p_min.z, b.p_min.z)};
        r.p_max = {rmax(p_max.x, b.p_max.x), rmax(p_max.y, b.p_max.y), rmax(p_max.z, b.p_max.z)};
        return r;
    }
    AABB union_vec(V3 v) const { // :786-791
        AABB r;
        r.p_min = {rmin(p_min.x, v.x), rmin(p_min.y, v.y), rmin(p_min.z, v.z)};
        r.p_max = {rmax(p_max.x, v.x), rmax(p_max.y, v.y), rmax(p_max.z, v.z)};
        return r;
    }
    V3 size() const { return p_max - p_min; }           // :839-841
    V3 center() const { return size() * 0.5f + p_min; } // :844-846
    float surface_area() const {                        // :822-836 (half the true area)
        V3 d = size();
        float s = 0.0f;
        for (int i = 0; i < 3; i++) {
            float v = 1.0f;
            for (int j = 0; j < 3; j++) {
                if (i == j) continue;
                v *= comp(d, j);
            }
            s += v;
        }
        return s;
    }
    bool intersect(const Ray &r, float *t_out) const { // :849-869
        float t_max = r.tfar, t_min = r.tnear;
        for (int d = 0; d < 3; d++) {
            float inv_d = 1.0f / comp(r.d, d);
            float t0 = (comp(p_min, d) - comp(r.o, d)) * inv_d;
            float t1 = (comp(p_max, d) - comp(r.o, d)) * inv_d;
            if (inv_d < 0.0f) std::swap(t0, t1);
            t_min = t0 > t_min ? t0 : t_min;
            t_max = t1 < t_max ? t1 : t_max;
            if (t_max <= t_min) return false;
        }
        *t_out = t_min;
        return true;
    }
};

// ============================================================================================
// math.rs
// ============================================================================================
struct P2 {
    float x, y;
};
P2 concentric_sample_disk(const Math &m, P2 u) { // :37-59
    P2 u_offset{u.x * 2.0f - 1.0f, u.y * 2.0f - 1.0f};
    if (u_offset.x == 0.0f && u_offset.y == 0.0f) return P2{0.0f, 0.0f};
    float theta, r;
    if (std::fabs(u_offset.x) > std::fabs(u_offset.y)) {
        r = u_offset.x;
        theta = FRAC_PI_4 * (u_offset.y / u_offset.x);
    } else {
        r = u_offset.y;
        theta = FRAC_PI_2 - FRAC_PI_4 * (u_offset.x / u_offset.y);
    }
    float s, c;
    m.sincos(theta, &s, &c);
    return P2{c * r, s * r};
}
V3 cosine_sample_hemisphere(const Math &m, P2 u) { // :61-65
    P2 d = concentric_sample_disk(m, u);
    float z = std::sqrt(rmax(0.0f, 1.0f - d.x * d.x - d.y * d.y));
    return V3{d.x, d.y, z};
}
struct Frame { // :357-384 (Matrix3 columns x, y, z)
    V3 x, y, z;
    explicit Frame(V3 n) {
        float sign = std::copysign(1.0f, n.z); // f32::signum: +1 for +0.0, -1 for -0.0 (NaN -> NaN, not reachable)
        float a = -1.0f / (sign + n.z);
        float b = n.x * n.y * a;
        x = V3{1.0f + sign * n.x * n.x * a, sign * b, -sign * n.x};
        y = V3{b, sign + n.y * n.y * a, -n.y};
        z = n;
    }
    Frame() : x{1, 0, 0}, y{0, 1, 0}, z{0, 0, 1} {}
    V3 to_world(V3 v) const { return x * v.x + y * v.y + z * v.z; }
    V3 to_local(V3 v) const { return V3{dot(v, x), dot(v, y), dot(v, z)}; }
};
P2 uniform_sample_triangle(P2 u) { // :388-394
    float su0 = std::sqrt(u.x);
    return P2{1.0f - su0, u.y * su0};
}
struct Distribution1D { // :402-406
    std::vector<float> cdf, func;
    float func_int = 0;
    static Distribution1D normalize(const std::vector<float> &elements) { // :418-441
        Distribution1D d;
        float cur = 0.0f;
        for (float e : elements) {
            d.cdf.push_back(cur);
            cur += e / (float)elements.size();
        }
        d.cdf.push_back(cur);
        if (cur != 0.0f)
            for (float &x : d.cdf) x /= cur;
        d.cdf.back() = 1.0f;
        d.func = elements;
        d.func_int = cur;
        return d;
    }
    // :447-457.  Rust's binary_search_by returns Ok(i) for cdf[i]==v, else Err(insertion)->insertion-1;
    // both equal "last index with cdf[i] <= v" when the cdf has no repeated entries.
    size_t sample_discrete(float v) const {
        size_t ub = (size_t)(std::upper_bound(cdf.begin(), cdf.end(), v) - cdf.begin());
        return ub - 1;
    }
    float pdf(size_t i) const { return cdf[i + 1] - cdf[i]; }            // :480-482
    float total() const { return func_int * (float)(cdf.size() - 1); }   // :484-486
};

// ============================================================================================
// bsdfs
// ============================================================================================
struct UV { // Option<Vector2<f32>>
    bool some = false;
    P2 v{0.0f, 0.0f};
};
struct BitmapTex { // structure.rs:382-386 (the part pixel_uv needs)
    uint32_t size_x = 0, size_y = 0;
    std::vector<Color> colors;
    Color pixel_uv(P2 uv) const { // :434-453
        auto modulo = [](float x, float n) { return std::fmod(std::fmod(x, n) + n, n); }; // tools.rs:39-41
        uv.x = modulo(uv.x, 1.0f);
        uv.y = modulo(uv.y, 1.0f);
        auto as_usize = [](float v) -> uint64_t { // `as usize`: saturating, NaN -> 0
            if (!(v > 0.0f)) return 0;
            if (v >= 18446744073709551616.0f) return ~(uint64_t)0;
            return (uint64_t)v;
        };
        uint64_t x = as_usize(uv.x * (float)size_x), y = as_usize(uv.y * (float)size_y);
        uint64_t i = (uint64_t)size_x * y + x;
        if (i >= colors.size()) return Color::zero();
        return colors[i];
    }
};
struct BSDFColor { // bsdfs/mod.rs:11-101
    enum Kind { Constant, Bitmap, Checkerbord, Grid } kind = Constant;
    Color c{1, 1, 1};
    std::shared_ptr<BitmapTex> img;
    Color color0{}, color1{};
    float line_width = 0.0f;
    P2 offset{0, 0}, scale{1, 1};
    static int32_t as_i32(float v) { // `as i32`: saturating, NaN -> 0
        if (v != v) return 0;
        if (v >= 2147483648.0f) return 2147483647;
        if (v <= -2147483648.0f) return (int32_t)0x80000000u;
        return (int32_t)v;
    }
    Color color(const UV &uv) const { // :32-101
        if (kind == Constant) return c;
        if (!uv.some) return Color::zero(); // "Found a texture but no uv coordinate given"
        if (kind == Bitmap) return img->pixel_uv(uv.v);
        if (kind == Checkerbord) {
            P2 q{uv.v.x * scale.x + offset.x, uv.v.y * scale.y + offset.y};
            int32_t x = 2 * (as_i32(q.x * 2.0f) % 2) - 1, y = 2 * (as_i32(q.y * 2.0f) % 2) - 1;
            return x * y == 1 ? color0 : color1;
        }
        P2 q{uv.v.x * scale.x + offset.x, (uv.v.y + scale.y) + offset.y}; // Grid: `uv.y + scale.y` (:84)
        float x = q.x - std::floor(q.x), y = q.y - std::floor(q.y);
        if (x > 0.5f) x -= 1.0f;
        if (y > 0.5f) y -= 1.0f;
        return (std::fabs(x) < line_width || std::fabs(y) < line_width) ? color0 : color1;
    }
};
struct SampledDirection { // bsdfs/mod.rs:129-137
    Color weight;
    V3 d;
    PDF pdf;
};
inline V3 reflect(V3 d) { return V3{-d.x, -d.y, d.z}; } // bsdfs/mod.rs:124-126

struct BSDF { // trait BSDF, bsdfs/mod.rs:163-199
    virtual ~BSDF() = default;
    virtual bool sample(const Math &m, const UV &uv, V3 d_in, P2 sample, SampledDirection *out) const = 0;
    virtual PDF pdf(const Math &m, const UV &uv, V3 d_in, V3 d_out) const = 0;
    virtual Color eval(const Math &m, const UV &uv, V3 d_in, V3 d_out) const = 0;
    virtual bool is_twosided() const = 0;
    virtual bool is_smooth() const = 0; // bsdf_type().is_smooth(), bsdfs/mod.rs:157-161
};
struct BSDFDiffuse : BSDF { // bsdfs/diffuse.rs
    BSDFColor diffuse;
    bool sample(const Math &m, const UV &uv, V3 d_in, P2 s, SampledDirection *out) const override { // :11-31
        if (d_in.z <= 0.0f) return false;
        V3 d_out = cosine_sample_hemisphere(m, s);
        *out = SampledDirection{diffuse.color(uv), d_out, PDF{PDF::SolidAngle, d_out.z * FRAC_1_PI}};
        return true;
    }
    PDF pdf(const Math &, const UV &uv, V3 d_in, V3 d_out) const override { // :33-51
        if (d_in.z <= 0.0f) return PDF{PDF::SolidAngle, 0.0f};
        if (d_out.z <= 0.0f) return PDF{PDF::SolidAngle, 0.0f};
        return PDF{PDF::SolidAngle, d_out.z * FRAC_1_PI};
    }
    Color eval(const Math &, const UV &uv, V3 d_in, V3 d_out) const override { // :53-71
        if (d_in.z <= 0.0f) return Color::zero();
        if (d_out.z > 0.0f) return diffuse.color(uv) * d_out.z * FRAC_1_PI;
        return Color::zero();
    }
    bool is_twosided() const override { return true; }
    bool is_smooth() const override { return false; }
};
struct BSDFPhong : BSDF { // bsdfs/phong.rs
    BSDFColor diffuse;
    BSDFColor specular;
    float exponent, weight_specular;
    bool sample(const Math &m, const UV &uv, V3 d_in, P2 s, SampledDirection *out) const override { // :14-63
        if (d_in.z <= 0.0f) return false;
        V3 d_out;
        if (s.x < weight_specular) {
            s.x /= weight_specular;
            float sin_alpha = std::sqrt(1.0f - m.powf(s.y, 2.0f / (exponent + 1.0f)));
            float cos_alpha = m.powf(s.y, 1.0f / (exponent + 1.0f));
            float phi = 2.0f * PI * s.x;
            float sp, cp;
            m.sincos(phi, &sp, &cp);
            V3 local_dir{sin_alpha * cp, sin_alpha * sp, cos_alpha};
            Frame frame(reflect(d_in));
            d_out = frame.to_world(local_dir);
            if (d_out.z <= 0.0f) return false;
        } else {
            s.x = (s.x - weight_specular) / (1.0f - weight_specular);
            d_out = cosine_sample_hemisphere(m, s);
        }
        PDF p = pdf(m, uv, d_in, d_out);
        if (p.value() == 0.0f) return false;
        *out = SampledDirection{eval(m, uv, d_in, d_out) / p.value(), d_out, p};
        return true;
    }
    PDF pdf(const Math &m, const UV &uv, V3 d_in, V3 d_out) const override { // :65-91
        if (d_in.z <= 0.0f || d_out.z <= 0.0f) return PDF{PDF::SolidAngle, 0.0f};
        float pdf_specular;
        float alpha = dot(reflect(d_in), d_out);
        if (alpha > 0.0f) pdf_specular = weight_specular * m.powf(alpha, exponent) * (exponent + 1.0f) / (2.0f * PI);
        else pdf_specular = 0.0f;
        float pdf_diffuse = (1.0f - weight_specular) * d_out.z * FRAC_1_PI;
        return PDF{PDF::SolidAngle, pdf_specular + pdf_diffuse};
    }
    Color eval(const Math &m, const UV &uv, V3 d_in, V3 d_out) const override { // :93-119
        if (d_in.z <= 0.0f || d_out.z <= 0.0f) return Color::zero();
        Color specular_value;
        float alpha = dot(reflect(d_in), d_out);
        if (alpha > 0.0f) specular_value = specular.color(uv) * (m.powf(alpha, exponent) * (exponent + 2.0f) / (2.0f * PI));
        else specular_value = Color::zero();
        Color diffuse_value = diffuse.color(uv) * d_out.z * FRAC_1_PI;
        return specular_value + diffuse_value;
    }
    bool is_twosided() const override { return true; }
    bool is_smooth() const override { return false; }
};
// f32::powi(n): llvm.powi / compiler-rt __powisf2 = binary exponentiation (r = 1; loop { if b&1 { r *= a } b /= 2; a *= a })
inline float powi(float a, int b) {
    float r = 1.0f;
    for (;;) {
        if (b & 1) r *= a;
        b /= 2;
        if (b == 0) break;
        a *= a;
    }
    return r;
}
// bsdfs/utils.rs
namespace bu {
inline float cos_theta(V3 w) { return w.z; }                                            // :5-7
inline float cos_2_theta(V3 w) { return w.z * w.z; }                                    // :8-10
inline float abs_cos_theta(V3 w) { return std::fabs(w.z); }                             // :11-13
inline float sin_2_theta(V3 w) { return rmax(1.0f - cos_2_theta(w), 0.0f); }            // :14-16
inline float sin_theta(V3 w) { return std::sqrt(sin_2_theta(w)); }                      // :17-19
inline float tan_theta(V3 w) { return sin_theta(w) / cos_theta(w); }                    // :20-22
inline float hypot2(float a, float b) {                                                 // :50-60
    if (std::fabs(a) > std::fabs(b)) {
        float r = b / a;
        return std::fabs(a) * std::sqrt(1.0f + r * r);
    } else if (b != 0.0f) {
        float r = a / b;
        return std::fabs(b) * std::sqrt(1.0f + r * r);
    } else return 0.0f;
}
inline V3 reflect_vector(V3 wo, V3 n) { return -(wo) + n * 2.0f * dot(wo, n); }          // :62-64
inline bool check_reflection_condition(V3 wi, V3 wo) {                                  // :65-67
    return std::fabs(wi.z * wo.z - wi.x * wo.x - wi.y * wo.y - 1.0f) < 0.0001f;
}
Color fresnel_conductor(float cos_theta, Color eta, Color k) {                          // :78-100
    float cos_theta_2 = cos_theta * cos_theta;
    float sin_theta_2 = 1.0f - cos_theta_2;
    float sin_theta_4 = sin_theta_2 * sin_theta_2;
    Color temp1 = eta * eta - k * k - color_value(sin_theta_2);
    Color a2pb2 = safe_sqrt(temp1 * temp1 + k * k * eta * eta * 4.0f);
    Color a = safe_sqrt((a2pb2 + temp1) * 0.5f);
    Color term1 = a2pb2 + color_value(cos_theta_2);
    Color term2 = a * (2.0f * cos_theta_2);
    Color rs2 = (term1 - term2) / (term1 + term2);
    Color term3 = a2pb2 * cos_theta_2 + color_value(sin_theta_4);
    Color term4 = term2 * sin_theta_2;
    Color rp2 = rs2 * (term3 - term4) / (term3 + term4);
    return 0.5f * (rp2 + rs2);
}
std::pair<float, float> fresnel_dielectric(float cos_theta_i_, float eta) {             // :103-130 -> (fresnel, cosThetaT)
    if (eta == 1.0f) return {0.0f, -cos_theta_i_};
    float scale = cos_theta_i_ > 0.0f ? 1.0f / eta : eta;
    float cos_theta_t_sqr = 1.0f - (1.0f - cos_theta_i_ * cos_theta_i_) * (scale * scale);
    if (cos_theta_t_sqr <= 0.0f) return {1.0f, 0.0f};
    float cos_theta_i = std::fabs(cos_theta_i_);
    float cos_theta_t = std::sqrt(cos_theta_t_sqr);
    float rs = (cos_theta_i - eta * cos_theta_t) / (cos_theta_i + eta * cos_theta_t);
    float rp = (eta * cos_theta_i - cos_theta_t) / (eta * cos_theta_i + cos_theta_t);
    float ct = cos_theta_i_ > 0.0f ? -cos_theta_t : cos_theta_t;
    return {0.5f * (rs * rs + rp * rp), ct};
}
} // namespace bu

// bsdfs/distribution.rs
struct MicrofacetDistribution {
    enum Type { Beckmann, GGX } microfacet_type;
    float alpha_u, alpha_v;
    float eval(const Math &mm, V3 m) const { // :27-56
        if (bu::cos_theta(m) <= 0.0f) return 0.0f;
        float cos_theta_2 = bu::cos_2_theta(m);
        float beckmann_exp = ((m.x * m.x) / (alpha_u * alpha_u) + (m.y * m.y) / (alpha_v * alpha_v)) / cos_theta_2;
        float res;
        if (microfacet_type == Beckmann) res = mm.exp(-beckmann_exp) / (PI * alpha_u * alpha_v * cos_theta_2 * cos_theta_2);
        else {
            float root = (1.0f + beckmann_exp) * cos_theta_2;
            res = 1.0f / (PI * alpha_u * alpha_v * root * root);
        }
        if (res * bu::cos_theta(m) < 1e-20f) return 0.0f;
        return res;
    }
    float pdf(const Math &mm, V3 m) const { return eval(mm, m) * bu::cos_theta(m); } // :58-60
    std::pair<V3, float> sample(const Math &mm, P2 sample) const { // :63-111 (asserts alpha_u == alpha_v)
        float sin_phi_m, cos_phi_m;
        mm.sincos(2.0f * PI * sample.y, &sin_phi_m, &cos_phi_m);
        float alpha_sqr = alpha_u * alpha_v;
        float cos_theta_m, pdf;
        if (microfacet_type == Beckmann) {
            float tan_theta_m_sqr = alpha_sqr * -mm.ln(1.0f - sample.x);
            cos_theta_m = 1.0f / std::sqrt(1.0f + tan_theta_m_sqr);
            pdf = (1.0f - sample.x) / (PI * alpha_u * alpha_v * powi(cos_theta_m, 3));
        } else {
            float tan_theta_m_sqr = alpha_sqr * sample.x / (1.0f - sample.x);
            cos_theta_m = 1.0f / std::sqrt(1.0f + tan_theta_m_sqr);
            float tmp = 1.0f + tan_theta_m_sqr / alpha_sqr;
            pdf = FRAC_1_PI / (alpha_u * alpha_v * powi(cos_theta_m, 3) * powi(tmp, 2));
        }
        if (pdf < 1e-20f) pdf = 0.0f;
        float sin_theta_m = std::sqrt(rmax(1.0f - powi(cos_theta_m, 2), 0.0f));
        return {V3{sin_theta_m * cos_phi_m, sin_theta_m * sin_phi_m, cos_theta_m}, pdf};
    }
    float smith_g1(V3 v, V3 m) const { // :117-144
        if (dot(v, m) * bu::cos_theta(v) <= 0.0f) return 0.0f;
        float tan_theta = std::fabs(bu::tan_theta(v));
        if (tan_theta == 0.0f) return 1.0f;
        float alpha = alpha_u;
        if (microfacet_type == Beckmann) {
            float a = 1.0f / (alpha * tan_theta);
            if (a >= 1.6f) return 1.0f;
            float a_sqr = powi(a, 2);
            return (3.535f * a + 2.181f * a_sqr) / (1.0f + 2.276f * a + 2.577f * a_sqr);
        }
        float root = alpha * tan_theta;
        return 2.0f / (1.0f + bu::hypot2(1.0f, root));
    }
    float g(V3 wi, V3 wo, V3 m) const { return smith_g1(wi, m) * smith_g1(wo, m); } // :113-115
};

struct BSDFMetal : BSDF { // bsdfs/metal.rs
    BSDFColor specular, eta, k;
    bool has_distribution;
    MicrofacetDistribution distr;
    bool sample(const Math &mm, const UV &uv, V3 d_in, P2 s, SampledDirection *out) const override { // :15-73
        if (d_in.z <= 0.0f) return false;
        if (!has_distribution) {
            *out = SampledDirection{specular.color(uv) * bu::fresnel_conductor(d_in.z, eta.color(uv), k.color(uv)), reflect(d_in), PDF{PDF::Discrete, 1.0f}};
            return true;
        }
        auto [m, pdf] = distr.sample(mm, s);
        if (pdf == 0.0f) return false;
        V3 wo = bu::reflect_vector(d_in, m);
        if (bu::cos_theta(wo) <= 0.0f) return false;
        Color f = bu::fresnel_conductor(dot(d_in, m), eta.color(uv), k.color(uv)) * specular.color(uv);
        float w = distr.eval(mm, m) * distr.g(d_in, wo, m) * dot(d_in, m) / (pdf * bu::cos_theta(d_in));
        *out = SampledDirection{w * f, wo, PDF{PDF::SolidAngle, pdf}};
        return true;
    }
    PDF pdf(const Math &mm, const UV &uv, V3 wi, V3 wo) const override { // :75-110 (Domain::SolidAngle; the Discrete arm is never asked)
        V3 h = normalize(wi + wo);
        return PDF{PDF::SolidAngle, distr.pdf(mm, h) / (4.0f * std::fabs(dot(wo, h)))};
    }
    Color eval(const Math &mm, const UV &uv, V3 wi, V3 wo) const override { // :112-156
        V3 h = normalize(wi + wo);
        float d = distr.eval(mm, h);
        if (d == 0.0f) return Color::zero();
        Color f = specular.color(uv) * bu::fresnel_conductor(dot(wi, h), eta.color(uv), k.color(uv));
        float g = distr.g(wi, wo, h);
        float model = d * g / (4.0f * bu::cos_theta(wi));
        return f * model;
    }
    bool is_twosided() const override { return true; }
    bool is_smooth() const override { return !has_distribution; } // DELTA without a distribution, GLOSSY with one (:166-171)
};
struct BSDFGlass : BSDF { // bsdfs/glass.rs
    BSDFColor specular_transmittance, specular_reflectance;
    float eta, inv_eta;
    V3 refract(V3 wi, float cos_theta_t) const { // :50-58
        float scale = cos_theta_t < 0.0f ? -inv_eta : -eta;
        return V3{scale * wi.x, scale * wi.y, cos_theta_t};
    }
    bool sample(const Math &, const UV &uv, V3 d_in, P2 s, SampledDirection *out) const override { // :75-121, transport == Importance
        auto [fresnel, cos_theta_trans] = bu::fresnel_dielectric(d_in.z, eta);
        if (s.x <= fresnel) {
            *out = SampledDirection{specular_reflectance.color(uv), reflect(d_in), PDF{PDF::Discrete, fresnel}};
        } else {
            float factor = 1.0f;
            *out = SampledDirection{specular_transmittance.color(uv) * factor * factor, refract(d_in, cos_theta_trans), PDF{PDF::Discrete, fresnel}};
        }
        return true;
    }
    // pdf() is todo!() and eval() asserts Domain::Discrete in the reference (:123-176): never reached because the BSDF is smooth
    PDF pdf(const Math &, const UV &, V3, V3) const override { std::abort(); }
    Color eval(const Math &, const UV &, V3, V3) const override { std::abort(); }
    bool is_twosided() const override { return false; }
    bool is_smooth() const override { return true; }
};
struct BSDFSubstrate : BSDF { // bsdfs/substrate.rs
    BSDFColor specular;
    BSDFColor diffuse;
    bool has_distribution;
    MicrofacetDistribution distr;
    Color schlick_fresnel(const UV &uv, float cos_theta) const { // :15-18
        Color rs = specular.color(uv);
        return rs + (Color::one() - rs) * powi(1.0f - cos_theta, 5);
    }
    PDF pdf_domain(const Math &mm, V3 wi, V3 wo, PDF::Kind domain) const { // :92-147
        if (wi.z <= 0.0f || wo.z <= 0.0f) return PDF{domain, 0.0f};
        V3 m = wi + wo;
        if (m.x == 0.0f && m.y == 0.0f && m.z == 0.0f) return PDF{domain, 0.0f};
        m = normalize(m);
        if (domain == PDF::Discrete) {
            if (bu::check_reflection_condition(wi, wo)) return PDF{PDF::Discrete, 0.5f};
            std::abort(); // unimplemented!()
        }
        float pdf_diffuse = wo.z * FRAC_1_PI;
        float pdf_specular = has_distribution ? distr.pdf(mm, m) / (4.0f * std::fabs(dot(wo, m))) : 0.0f;
        return PDF{PDF::SolidAngle, 0.5f * (pdf_diffuse + pdf_specular)};
    }
    Color eval_domain(const Math &mm, const UV &uv, V3 d_in, V3 d_out, PDF::Kind domain) const { // :149-206
        if (d_in.z <= 0.0f || d_out.z <= 0.0f) return Color::zero();
        V3 m = d_in + d_out;
        if (m.x == 0.0f && m.y == 0.0f && m.z == 0.0f) return Color::zero();
        m = normalize(m);
        if (domain == PDF::SolidAngle) {
            Color diff = diffuse.color(uv) * (Color::one() - specular.color(uv)) * (28.0f / (23.0f * PI)) * (1.0f - powi(1.0f - 0.5f * bu::abs_cos_theta(d_in), 5)) *
                         (1.0f - powi(1.0f - 0.5f * bu::abs_cos_theta(d_out), 5));
            Color spec = Color::zero();
            if (has_distribution) {
                float model = distr.eval(mm, m) / (4.0f * std::fabs(dot(d_in, m)) * rmax(std::fabs(bu::cos_theta(d_in)), std::fabs(bu::cos_theta(d_out))));
                spec = model * schlick_fresnel(uv, dot(d_in, m));
            }
            return (diff + spec) * d_out.z;
        }
        if (bu::check_reflection_condition(d_in, d_out)) return schlick_fresnel(uv, dot(d_in, m));
        std::abort(); // unimplemented!()
    }
    bool sample(const Math &mm, const UV &uv, V3 d_in, P2 s, SampledDirection *out) const override { // :22-90
        if (d_in.z <= 0.0f) return false;
        V3 d_out;
        PDF::Kind domain;
        if (s.x < 0.5f) {
            s.x *= 2.0f;
            d_out = cosine_sample_hemisphere(mm, s);
            domain = PDF::SolidAngle;
        } else {
            s.x = (s.x - 0.5f) * 2.0f;
            V3 m;
            if (!has_distribution) {
                m = V3{0.0f, 0.0f, 1.0f};
                domain = PDF::Discrete;
            } else {
                auto r = distr.sample(mm, s);
                if (r.second == 0.0f) return false;
                m = r.first;
                domain = PDF::SolidAngle;
            }
            d_out = bu::reflect_vector(d_in, m);
            if (bu::cos_theta(d_out) <= 0.0f) return false;
        }
        PDF p = pdf_domain(mm, d_in, d_out, domain);
        if (p.value() == 0.0f) return false;
        Color f = eval_domain(mm, uv, d_in, d_out, domain);
        *out = SampledDirection{f / p.value(), d_out, p};
        return true;
    }
    PDF pdf(const Math &mm, const UV &uv, V3 wi, V3 wo) const override { return pdf_domain(mm, wi, wo, PDF::SolidAngle); }
    Color eval(const Math &mm, const UV &uv, V3 wi, V3 wo) const override { return eval_domain(mm, uv, wi, wo, PDF::SolidAngle); }
    bool is_twosided() const override { return true; }
    bool is_smooth() const override { return !has_distribution; } // DELTA | DIFFUSE without a distribution (:216-221)
};
struct BSDFBlend : BSDF { // bsdfs/blend.rs
    std::unique_ptr<BSDF> bsdf1, bsdf2;
    float weight;
    bool sample(const Math &m, const UV &uv, V3 d_in, P2 s, SampledDirection *out) const override { // :10-45
        // assert!(!bsdf1.is_smooth() && !bsdf2.is_smooth()) (:17) is enforced when the scene is built
        SampledDirection sd;
        bool some;
        if (s.x < weight) {
            P2 scaled{s.x * (1.0f / weight), s.y};
            some = bsdf1->sample(m, uv, d_in, scaled, &sd);
        } else {
            P2 scaled{(s.x - weight) * (1.0f / (1.0f - weight)), s.y};
            some = bsdf2->sample(m, uv, d_in, scaled, &sd);
        }
        if (!some) return false;
        sd.pdf = pdf(m, uv, d_in, sd.d);
        if (sd.pdf.value() == 0.0f) return false;
        sd.weight = eval(m, uv, d_in, sd.d) / sd.pdf.value();
        *out = sd;
        return true;
    }
    PDF pdf(const Math &m, const UV &uv, V3 d_in, V3 d_out) const override { // :47-62 (`PDF * f32`, structure.rs:85-94)
        PDF pdf_1 = bsdf1->pdf(m, uv, d_in, d_out), pdf_2 = bsdf2->pdf(m, uv, d_in, d_out);
        pdf_1.v = pdf_1.v * weight, pdf_2.v = pdf_2.v * (1.0f - weight);
        if (pdf_1.kind != PDF::SolidAngle || pdf_2.kind != PDF::SolidAngle) throw std::runtime_error("get wrong type of BSDF");
        return PDF{PDF::SolidAngle, pdf_1.v + pdf_2.v};
    }
    Color eval(const Math &m, const UV &uv, V3 d_in, V3 d_out) const override { // :64-76 (`f32 * Color`, structure.rs:294-303: plain products)
        Color a = bsdf1->eval(m, uv, d_in, d_out), b = bsdf2->eval(m, uv, d_in, d_out);
        const float w2 = 1.0f - weight;
        return Color{a.r * weight, a.g * weight, a.b * weight} + Color{b.r * w2, b.g * w2, b.b * w2};
    }
    bool is_twosided() const override { return true; } // :84-89 (panics unless both parts are two-sided; all rough BSDFs are)
    bool is_smooth() const override { return bsdf1->is_smooth() || bsdf2->is_smooth(); } // bsdf_type() = bsdf1 | bsdf2 (:91-93)
};
std::unique_ptr<BSDF> make_bsdf(const rl_material &m, const rl_texture *textures = nullptr, uint32_t ntextures = 0, const rl_material *subs = nullptr,
                                uint32_t nsubs = 0) {
    if (m.kind == RL_BSDF_BLEND) {
        if (!subs || m.blend_a == 0 || m.blend_b == 0 || m.blend_a > nsubs || m.blend_b > nsubs || subs[m.blend_a - 1].kind == RL_BSDF_BLEND ||
            subs[m.blend_b - 1].kind == RL_BSDF_BLEND)
            throw std::runtime_error("blend: bad submaterial reference");
        auto b = std::make_unique<BSDFBlend>();
        b->bsdf1 = make_bsdf(subs[m.blend_a - 1], textures, ntextures), b->bsdf2 = make_bsdf(subs[m.blend_b - 1], textures, ntextures);
        b->weight = m.blend_weight;
        if (b->bsdf1->is_smooth() || b->bsdf2->is_smooth()) throw std::runtime_error("blend: smooth part (blend.rs:17)");
        return b;
    }
    // a colour slot: BSDFColor::Constant(rgb) or one of the scene's textures (bsdf_texture_match_pbrt, bsdfs/mod.rs:218-240)
    auto slot = [&](const float *rgb, uint32_t tex) {
        BSDFColor c;
        c.c = Color{rgb[0], rgb[1], rgb[2]};
        if (tex != 0 && textures && tex <= ntextures) {
            const rl_texture &t = textures[tex - 1];
            c.kind = t.kind == RL_TEX_BITMAP ? BSDFColor::Bitmap : (t.kind == RL_TEX_GRID ? BSDFColor::Grid : BSDFColor::Checkerbord);
            c.color0 = Color{t.color0[0], t.color0[1], t.color0[2]}, c.color1 = Color{t.color1[0], t.color1[1], t.color1[2]};
            c.line_width = t.line_width, c.offset = P2{t.offset[0], t.offset[1]}, c.scale = P2{t.scale[0], t.scale[1]};
            if (t.kind == RL_TEX_BITMAP) {
                c.img = std::make_shared<BitmapTex>();
                c.img->size_x = t.width, c.img->size_y = t.height;
                for (size_t i = 0; i < (size_t)t.width * t.height; i++) c.img->colors.push_back(Color{t.pixels[3 * i], t.pixels[3 * i + 1], t.pixels[3 * i + 2]});
            }
        }
        return c;
    };
    auto kd_color = [&]() { return slot(m.kd, m.kd_texture); };
    auto distribution = [&](bool *has) {
        *has = m.microfacet != RL_MICROFACET_NONE;
        return MicrofacetDistribution{m.microfacet == RL_MICROFACET_BECKMANN ? MicrofacetDistribution::Beckmann : MicrofacetDistribution::GGX, m.alpha, m.alpha};
    };
    if (m.kind == RL_BSDF_METAL) {
        auto b = std::make_unique<BSDFMetal>();
        b->specular = slot(m.ks, m.ks_texture), b->eta = slot(m.eta, m.eta_texture), b->k = slot(m.k, m.k_texture);
        b->distr = distribution(&b->has_distribution);
        return b;
    }
    if (m.kind == RL_BSDF_GLASS) {
        auto b = std::make_unique<BSDFGlass>();
        b->specular_reflectance = slot(m.ks, m.ks_texture), b->specular_transmittance = slot(m.kt, m.kt_texture);
        b->eta = m.ior;
        b->inv_eta = 1.0f / b->eta; // glass.rs:46
        return b;
    }
    if (m.kind == RL_BSDF_SUBSTRATE) {
        auto b = std::make_unique<BSDFSubstrate>();
        b->diffuse = kd_color(), b->specular = slot(m.ks, m.ks_texture);
        b->distr = distribution(&b->has_distribution);
        return b;
    }
    if (m.kind == RL_BSDF_PHONG) {
        auto b = std::make_unique<BSDFPhong>();
        b->diffuse = kd_color();
        b->specular = slot(m.ks, m.ks_texture);
        b->exponent = m.exponent;
        b->weight_specular = m.weight_specular;
        return b;
    }
    auto b = std::make_unique<BSDFDiffuse>();
    b->diffuse = kd_color();
    return b;
}

// ============================================================================================
// geometry.rs: Mesh
// ============================================================================================
struct IntersectionUV { // structure.rs:924-930
    float t;
    V3 p, n;
    float u, v;
};
struct SampledPosition { // structure.rs:96-102
    V3 p, n;
    PDF pdf;
    size_t primitive_id;
    UV uv{}; // Option<Vector2<f32>>
};
struct Idx3 {
    uint32_t x, y, z;
};
struct Mesh { // geometry.rs:107-119
    std::vector<V3> vertices;
    std::vector<Idx3> indices;
    bool has_normals = false;
    std::vector<V3> normals;
    bool has_uv = false;
    std::vector<P2> uv; // Mesh.uv: Option<Vec<Vector2<f32>>>
    std::unique_ptr<BSDF> bsdf;
    bool light = false; // emission != EmissionType::Zero
    Color emission = Color::zero();
    // EmissionType (:99-104): Color { v } | HSV { scale } | Texture { scale, img }
    enum EmissionType { EColor = 1, EHsv = 2, ETexture = 3 } emission_type = EColor;
    float emission_scale = 0.0f;
    std::shared_ptr<BitmapTex> emission_img;
    Distribution1D cdf;
    uint32_t first_prim = 0; // global index of triangle 0 (mesh-major numbering)

    void build_cdf() { // Mesh::new, :130-138
        std::vector<float> areas;
        for (auto id : indices) {
            V3 v0 = vertices[id.x], v1 = vertices[id.y], v2 = vertices[id.z];
            areas.push_back(magnitude(cross(v1 - v0, v2 - v0)) * 0.5f);
        }
        cdf = Distribution1D::normalize(areas);
    }
    Color emit(const UV &uv) const { // :184-206
        if (!light) return Color::zero();
        if (emission_type == EHsv) {
            Color c1{1.0f, 0.0f, 0.0f}, c2{0.0f, 1.0f, 0.0f};
            if (!uv.some) std::abort(); // uv.unwrap()
            float x = std::fmod(std::fabs(uv.v.x), 1.0f); // uv.x.abs() % 1.0
            Color c = x * c1 + (1.0f - x) * c2;
            return c * emission_scale;
        }
        if (emission_type == ETexture) {
            if (!uv.some) std::abort();
            return emission_img->pixel_uv(uv.v) * emission_scale;
        }
        return emission;
    }
    bool is_light() const { return light; }                         // :412-417
    float pdf() const { return 1.0f / cdf.total(); }                // :223-225

    bool intersection_tri(size_t i, V3 p_c, V3 d_c, IntersectionUV &its) const { // :358-410
        Idx3 id = indices[i];
        V3 v0 = vertices[id.x], v1 = vertices[id.y], v2 = vertices[id.z];
        V3 e1 = v1 - v0, e2 = v2 - v0;
        V3 n_geo = normalize(cross(e1, e2));
        float denom = dot(d_c, n_geo);
        if (denom == 0.0f) return false;
        float t = -dot(p_c - v0, n_geo) / denom;
        if (t < 0.0f) return false;
        V3 p = p_c + t * d_c;
        float det = magnitude(cross(e1, e2));
        V3 u0 = cross(e1, p - v0);
        V3 v0c = cross(p - v0, e2);
        if (dot(u0, n_geo) < 0.0f || dot(v0c, n_geo) < 0.0f) return false;
        float v = magnitude(u0) / det;
        float u = magnitude(v0c) / det;
        if (u < 0.0f || v < 0.0f || u > 1.0f || v > 1.0f) return false;
        if (u + v <= 1.0f) {
            if (t < its.t && t > 0.00001f) {
                its.t = t, its.u = u, its.v = v, its.p = p, its.n = n_geo;
                return true;
            }
        }
        return false;
    }
    AABB compute_aabb_tri(size_t i) const { // :423-439
        Idx3 id = indices[i];
        AABB aabb;
        aabb = aabb.union_vec(vertices[id.x]);
        aabb = aabb.union_vec(vertices[id.y]);
        aabb = aabb.union_vec(vertices[id.z]);
        V3 s = aabb.size();
        if (s.x < EPSILON) aabb.p_max.x += EPSILON, aabb.p_min.x -= EPSILON;
        if (s.y < EPSILON) aabb.p_max.y += EPSILON, aabb.p_min.y -= EPSILON;
        if (s.z < EPSILON) aabb.p_max.z += EPSILON, aabb.p_min.z -= EPSILON;
        return aabb;
    }
    SampledPosition sample_tri(size_t primitive_id, P2 v) const { // :261-337
        Idx3 id = indices[primitive_id];
        V3 v0 = vertices[id.x], v1 = vertices[id.y], v2 = vertices[id.z];
        P2 b = uniform_sample_triangle(v);
        V3 pos = v0 * b.x + v1 * b.y + v2 * (1.0f - b.x - b.y);
        V3 n_g;
        {
            V3 u = v1 - v0, w = v2 - v0;
            n_g = normalize(cross(w, u));
        }
        if (has_normals) {
            V3 n0 = normals[id.x], n1 = normals[id.y], n2 = normals[id.z];
            V3 n = n0 * b.x + n1 * b.y + n2 * (1.0f - b.x - b.y);
            float n_l = magnitude2(n);
            if (n_l == 0.0f) n = n_g;
            else if (n_l != 1.0f) n = n / std::sqrt(n_l);
            if (dot(n_g, n) < 0.0f) n_g = -n_g;
        }
        UV suv;
        if (has_uv) { // :316-325: the interpolated uv, normalized as a 2-vector (sic)
            P2 n0 = uv[id.x], n1 = uv[id.y], n2 = uv[id.z];
            float b2 = 1.0f - b.x - b.y;
            P2 q{n0.x * b.x + n1.x * b.y + n2.x * b2, n0.y * b.x + n1.y * b.y + n2.y * b2};
            float il = 1.0f / std::sqrt(q.x * q.x + q.y * q.y); // normalize = v * (1 / |v|)
            suv.some = true, suv.v = P2{q.x * il, q.y * il};
        }
        float area_tri = magnitude(cross(v1 - v0, v2 - v0)) * 0.5f;
        return SampledPosition{pos, n_g, PDF{PDF::Area, 1.0f / area_tri}, primitive_id, suv};
    }
    float pdf_tri(size_t primitive_id) const { // :226-234
        Idx3 id = indices[primitive_id];
        V3 v0 = vertices[id.x], v1 = vertices[id.y], v2 = vertices[id.z];
        float area_tri = magnitude(cross(v1 - v0, v2 - v0)) * 0.5f;
        return 1.0f / area_tri;
    }
    SampledPosition sample(float s, P2 v) const { // :340-348
        size_t primitive_id = cdf.sample_discrete(s);
        SampledPosition res = sample_tri(primitive_id, v);
        res.pdf = PDF{PDF::Area, 1.0f / cdf.total()};
        return res;
    }
    Color flux() const { // emitter.rs:591-599 (f32*Color then Color*f32); HSV / Texture: Color::value(scale), "TODO" there
        Color e = !light ? Color::zero() : (emission_type == EColor ? emission : Color{emission_scale, emission_scale, emission_scale});
        return cdf.total() * e * PI;
    }
};

// ============================================================================================
// structure.rs: Intersection::fill_intersection :965-1059
// ============================================================================================
struct Intersection {
    float dist;
    V3 n_g, n_s, p;
    const Mesh *mesh;
    Frame frame;
    V3 wi;
    size_t primitive_id;
    UV uv; // structure.rs:1015-1023
    float cos_theta() const { return wi.z; }
};
Intersection fill_intersection(const Mesh *mesh, size_t tri_id, float hit_u, float hit_v, const Ray &ray, V3 n_g, float dist, V3 p) {
    Idx3 index = mesh->indices[tri_id];
    V3 n_s;
    if (mesh->has_normals) {
        V3 d0 = mesh->normals[index.x], d1 = mesh->normals[index.y], d2 = mesh->normals[index.z];
        V3 ns = d0 * (1.0f - hit_u - hit_v) + d1 * hit_u + d2 * hit_v;
        if (dot(n_g, ns) < 0.0f) n_g = -n_g;
        float l = dot(ns, ns);
        if (l == 0.0f) n_s = n_g;
        else if (l != 1.0f) n_s = ns / std::sqrt(l);
        else n_s = ns;
    } else n_s = n_g;
    if (mesh->bsdf->is_twosided() && !mesh->is_light() && dot(ray.d, n_s) > 0.0f) {
        n_s = V3{-n_s.x, -n_s.y, -n_s.z};
        n_g = V3{-n_g.x, -n_g.y, -n_g.z};
    }
    Intersection its;
    its.dist = dist, its.n_g = n_g, its.n_s = n_s, its.p = p, its.mesh = mesh;
    its.frame = Frame(n_s);
    its.wi = its.frame.to_local(-ray.d);
    its.primitive_id = tri_id;
    if (mesh->has_uv) { // UV interpolation, structure.rs:1015-1023
        P2 d0 = mesh->uv[index.x], d1 = mesh->uv[index.y], d2 = mesh->uv[index.z];
        float w = 1.0f - hit_u - hit_v;
        its.uv.some = true;
        its.uv.v = P2{d0.x * w + d1.x * hit_u + d2.x * hit_v, d0.y * w + d1.y * hit_u + d2.y * hit_v};
    }
    return its;
}
inline Ray spawn_ray(const Intersection &its, V3 d_out) { return Ray{its.p, d_out, EPSILON, F32_MAX}; } // structure.rs:717-731

// ============================================================================================
// emitter.rs: impl Emitter for Mesh :570-688, EmitterSampler :1491-1647
// ============================================================================================
struct Emitter;
struct LightSampling { // :10-24
    const Emitter *emitter;
    PDF pdf;
    V3 p, n, d;
    size_t primitive_id;
    Color weight;
    UV uv{}; // uv of the sampled position (mesh emitters)
    bool is_valid() const { return !pdf.is_zero(); }
};
struct LightSamplingPDF { // :26-44
    V3 o, p, n, dir;
    uint32_t math_mode = ORC_MATH_SPEC; // (not in the reference: which sin / atan2 / acos EnvironmentLightColor::pdf evaluates with)
};
struct BoundingSphere { // structure.rs:880-884
    V3 center;
    float radius;
};
struct Emitter { // trait Emitter, emitter.rs:46-94 (the methods this path calls)
    virtual ~Emitter() = default;
    virtual bool is_surface() const { return false; }                                                    // :70-72
    virtual PDF direct_pdf_tri(const LightSamplingPDF &, size_t) const { std::abort(); }                 // unimplemented!() (:84-86)
    virtual LightSampling direct_sample_tri(V3, size_t, P2) const { std::abort(); }                      // unimplemented!() (:76-83)
    virtual PDF direct_pdf(const LightSamplingPDF &ls) const = 0;
    virtual LightSampling direct_sample(const Math &m, V3 p, float r, P2 uv) const = 0;
    virtual Color flux() const = 0;
    virtual Color eval(const UV &uv) const = 0; // eval(d, uv)
};
struct MeshEmitter : Emitter { // impl Emitter for Mesh, emitter.rs:570-688
    const Mesh *m;
    explicit MeshEmitter(const Mesh *mesh) : m(mesh) {}
    PDF direct_pdf(const LightSamplingPDF &ls) const override { // :571-579
        float cos_light = rmax(dot(ls.n, -ls.dir), 0.0f);
        if (cos_light == 0.0f) return PDF{PDF::SolidAngle, 0.0f};
        float geom = cos_light / magnitude2(ls.p - ls.o);
        return PDF{PDF::SolidAngle, m->pdf() / geom};
    }
    LightSampling direct_sample(const Math &, V3 p, float r, P2 uv) const override { // :652-688
        SampledPosition sp = m->sample(r, uv);
        V3 d = sp.p - p;
        float dist = magnitude(d);
        if (dist != 0.0f) d = d / dist;
        float geom = dist != 0.0f ? rmax(dot(sp.n, -d), 0.0f) / (dist * dist) : 0.0f;
        float pdf_area = sp.pdf.value();
        PDF pdf = sp.pdf.as_solid_angle_geom(geom);
        Color weight = pdf.is_zero() ? Color::zero() : m->emit(sp.uv) * geom / pdf_area;
        return LightSampling{this, pdf, sp.p, sp.n, d, sp.primitive_id, weight, sp.uv};
    }
    Color flux() const override { return m->flux(); }
    Color eval(const UV &uv) const override { return m->emit(uv); }
    bool is_surface() const override { return true; } // :727-729
    PDF direct_pdf_tri(const LightSamplingPDF &ls, size_t id_primitive) const override { // :581-589
        float cos_light = rmax(dot(ls.n, -ls.dir), 0.0f);
        if (cos_light == 0.0f) return PDF{PDF::SolidAngle, 0.0f};
        float geom = cos_light / magnitude2(ls.p - ls.o);
        return PDF{PDF::SolidAngle, m->pdf_tri(id_primitive) / geom};
    }
    LightSampling direct_sample_tri(V3 p, size_t primitive_id, P2 uv) const override { // :608-649
        SampledPosition sp = m->sample_tri(primitive_id, uv);
        V3 d = sp.p - p;
        float dist = magnitude(d);
        if (dist != 0.0f) d = d / dist;
        float geom = dist != 0.0f ? rmax(dot(sp.n, -d), 0.0f) / (dist * dist) : 0.0f;
        float pdf_area = sp.pdf.value();
        PDF pdf = sp.pdf.as_solid_angle_geom(geom);
        Color weight = pdf.is_zero() ? Color::zero() : m->emit(sp.uv) * geom / pdf_area;
        return LightSampling{this, pdf, sp.p, sp.n, d, 0, weight, sp.uv}; // primitive_id: None ("Not sampled a particular primitive")
    }
};
struct PointEmitter : Emitter { // emitter.rs:186-250
    Color intensity;
    V3 position;
    PDF direct_pdf(const LightSamplingPDF &) const override { return PDF{PDF::Discrete, 1.0f}; }
    LightSampling direct_sample(const Math &, V3 v, float, P2) const override { // :197-215
        V3 p = position;
        V3 d = p - v;
        float dist = magnitude(d);
        d = d / dist;
        return LightSampling{this, PDF{PDF::Discrete, 1.0f}, p, V3{0.0f, 0.0f, 0.0f}, d, 0, intensity / powi(dist, 2)};
    }
    Color flux() const override { return intensity * 4.0f * PI; } // :239-241
    Color eval(const UV &) const override { return intensity; }
};
struct DirectionalLight : Emitter { // emitter.rs:96-190
    V3 direction; // from the light to the world
    Color intensity;
    BoundingSphere bsphere{}; // preprocess(): scene.bsphere with radius * 1.1 (:106-109)
    PDF direct_pdf(const LightSamplingPDF &) const override { return PDF{PDF::Discrete, 1.0f}; }
    LightSampling direct_sample(const Math &, V3 v, float, P2) const override { // :115-133
        V3 p = v - bsphere.radius * direction;
        return LightSampling{this, PDF{PDF::Discrete, 1.0f}, p, direction, -direction, 0, intensity};
    }
    Color flux() const override { // :164-168
        float area = PI * powi(bsphere.radius, 2);
        return area * intensity;
    }
    Color eval(const UV &) const override { return intensity; }
};
// math.rs:324-352
inline bool solve_quadratic(float a, float b, float c, float *x0_out, float *x1_out) {
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
    float d_sqrt = std::sqrt(d);
    float tmp = b < 0.0f ? -0.5f * (b - d_sqrt) : -0.5f * (b + d_sqrt);
    float x0 = tmp / a, x1 = c / tmp;
    if (x0 > x1) *x0_out = x1, *x1_out = x0;
    else *x0_out = x0, *x1_out = x1;
    return true;
}
inline bool bsphere_intersect(const BoundingSphere &bs, const Ray &r, float *t) { // structure.rs:899-920 (b = +2 d_p.d, verbatim)
    V3 d_p = bs.center - r.o;
    float a = magnitude2(r.d);
    float b = 2.0f * dot(d_p, r.d);
    float c = magnitude2(d_p) - bs.radius * bs.radius;
    float t0, t1;
    if (!solve_quadratic(a, b, c, &t0, &t1)) return false;
    if (t0 < r.tnear) {
        if (t1 < r.tfar) {
            *t = t1;
            return true;
        }
        return false;
    } else if (t0 < r.tfar) {
        *t = t0;
        return true;
    }
    return false;
}
inline V3 sample_uniform_sphere(const Math &m, P2 u) { // math.rs:67-72
    float z = 1.0f - 2.0f * u.x;
    float r = std::sqrt(rmax(1.0f - z * z, 0.0f));
    float phi = 2.0f * PI * u.y;
    float sp, cp;
    m.sincos(phi, &sp, &cp);
    return V3{r * cp, r * sp, z};
}
constexpr float ONE_MINUS_EPSILON = 0.9999999403953552f; // lib.rs:52
inline float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); } // lib.rs:59-67
inline uint64_t as_usize(float v) { // `as usize` / `as u32` on the values used here: saturating, NaN -> 0
    if (!(v > 0.0f)) return 0;
    if (v >= 18446744073709551616.0f) return ~(uint64_t)0;
    return (uint64_t)v;
}
inline float dist1d_sample_continuous(const Distribution1D &d, float v) { // math.rs:459-478
    size_t i = d.sample_discrete(v);
    float dv = v - d.cdf[i];
    float pdf = d.pdf(i);
    if (pdf > 0.0f) dv = dv / pdf;
    return (float)i + dv;
}
struct Distribution2D { // math.rs:489-532
    Distribution1D marginal;
    std::vector<Distribution1D> conditionals;
    static Distribution2D from_bitmap(const BitmapTex &image) { // :495-521
        Distribution2D d;
        std::vector<float> marg;
        for (uint32_t y = 0; y < image.size_y; y++) {
            std::vector<float> cond;
            for (uint32_t x = 0; x < image.size_x; x++) {
                const Color p = image.colors[(size_t)y * image.size_x + x];
                cond.push_back(p.r * 0.212671f + p.g * 0.715160f + p.b * 0.072169f); // Color::luminance, structure.rs:173-176
            }
            d.conditionals.push_back(Distribution1D::normalize(cond));
            marg.push_back(d.conditionals.back().func_int);
        }
        d.marginal = Distribution1D::normalize(marg);
        return d;
    }
    P2 sample_continuous(P2 uv) const { // :523-527
        float y = dist1d_sample_continuous(marginal, uv.y);
        float x = dist1d_sample_continuous(conditionals[as_usize(y)], uv.x);
        return P2{x, y};
    }
    float pdf(size_t x, size_t y) const { return conditionals[y].func[x] / marginal.func_int; } // :529-531
};
inline P2 to_spherical_coordinates(const Math &m, V3 d) { // emitter.rs:320-338
    float p = m.atan2(d.y, d.x);
    if (p < 0.0f) p = p + 2.0f * PI;
    P2 uv{p * FRAC_1_PI * 0.5f, m.acos(clampf(d.z, -1.0f, 1.0f)) * FRAC_1_PI};
    uv.x = clampf(uv.x, 0.0f, ONE_MINUS_EPSILON);
    uv.y = clampf(uv.y, 0.0f, ONE_MINUS_EPSILON);
    return uv;
}
struct EnvironmentLightColor { // emitter.rs:300-427
    bool is_texture = false;
    Color constant{};
    BitmapTex image;
    Distribution2D image_cdf[2]; // [math mode]: new_texture weights the rows by sin((y + 0.5) PI / size.y) (:341-353)
    void new_texture(const BitmapTex &img) {
        is_texture = true;
        image = img;
        for (uint32_t mode = 0; mode < 2; mode++) {
            BitmapTex image_pdf = img;
            for (uint32_t y = 0; y < img.size_y; y++) {
                float w = Math{mode}.sin(((float)y + 0.5f) * PI / (float)img.size_y);
                for (uint32_t x = 0; x < img.size_x; x++) {
                    Color &c = image_pdf.colors[(size_t)y * img.size_x + x];
                    c.r *= w, c.g *= w, c.b *= w; // MulAssign<f32>, structure.rs:225-231
                }
            }
            image_cdf[mode] = Distribution2D::from_bitmap(image_pdf);
        }
    }
    void sample_direction(const Math &m, P2 uv, V3 *d, Color *color, float *pdf) const { // :354-391
        if (!is_texture) {
            *d = sample_uniform_sphere(m, uv), *color = constant, *pdf = 1.0f / (PI * 4.0f);
            return;
        }
        const Distribution2D &cdf = image_cdf[m.mode == ORC_MATH_LIBM ? ORC_MATH_LIBM : ORC_MATH_SPEC];
        P2 q = cdf.sample_continuous(uv);
        q.x = clampf(q.x, 0.0f, (float)image.size_x - 1.0f);
        q.y = clampf(q.y, 0.0f, (float)image.size_y - 1.0f);
        Color value = image.colors[(size_t)as_usize(q.y) * image.size_x + (size_t)as_usize(q.x)];
        float p = cdf.pdf((size_t)as_usize(q.x), (size_t)as_usize(q.y));
        float sin_phi, cos_phi, sin_theta, cos_theta;
        m.sincos((2.0f * PI / (float)image.size_x) * q.x, &sin_phi, &cos_phi);
        m.sincos((PI / (float)image.size_y) * q.y, &sin_theta, &cos_theta);
        *d = V3{sin_theta * cos_phi, sin_theta * sin_phi, cos_theta};
        if (sin_theta == 0.0f) *color = Color::zero(), *pdf = 0.0f;
        else *color = value, *pdf = p / (2.0f * (PI * PI) * sin_theta);
    }
    Color eval(const Math &m, V3 d) const { // :393-401
        if (!is_texture) return constant;
        return image.pixel_uv(to_spherical_coordinates(m, d));
    }
    float pdf(const Math &m, V3 d) const { // :403-425
        if (!is_texture) return 1.0f / (PI * 4.0f);
        P2 uv = to_spherical_coordinates(m, d);
        const Distribution2D &cdf = image_cdf[m.mode == ORC_MATH_LIBM ? ORC_MATH_LIBM : ORC_MATH_SPEC];
        float p = cdf.pdf((size_t)as_usize(uv.x * (float)image.size_x), (size_t)as_usize(uv.y * (float)image.size_y));
        float sin_theta = m.sin(PI * uv.y);
        if (sin_theta == 0.0f) return 0.0f;
        return p / (2.0f * (PI * PI) * sin_theta);
    }
};
struct EnvironmentLight : Emitter { // emitter.rs:428-568
    EnvironmentLightColor luminance;
    BoundingSphere bsphere{}; // preprocess(): scene.bsphere, radius * 1.1
    PDF direct_pdf(const LightSamplingPDF &ls) const override { return PDF{PDF::SolidAngle, luminance.pdf(Math{ls.math_mode}, ls.dir)}; } // :470-473
    LightSampling direct_sample(const Math &math, V3 v, float, P2 uv) const override { // :474-511
        V3 d;
        Color color;
        float pdf;
        luminance.sample_direction(math, uv, &d, &color, &pdf);
        float t;
        if (!bsphere_intersect(bsphere, ray_new(v, d), &t)) return LightSampling{this, PDF{PDF::SolidAngle, pdf}, V3{0, 0, 0}, V3{0, 0, 0}, d, 0, Color::zero()};
        V3 p = v + d * t;
        V3 n = normalize(bsphere.center - p);
        return LightSampling{this, PDF{PDF::SolidAngle, pdf}, p, n, d, 0, color / pdf};
    }
    Color flux() const override { // :512-524
        if (!luminance.is_texture) return (PI * powi(bsphere.radius, 2)) * luminance.constant;
        float v = PI * powi(bsphere.radius, 2) * luminance.image_cdf[ORC_MATH_SPEC].marginal.func_int; // (the SPEC table: what the device builds)
        return Color{v, v, v};
    }
    Color eval(const UV &) const override { return luminance.constant; }
};
// ---- the light tree of `-x ats`: emitter.rs:782-1400 ------------------------------------------------------------------------
inline float safe_acos(float v) { return std::acos(rmin(rmax(v, -1.0f), 1.0f)); } // :789-791
inline float safe_asin(float v) { return std::asin(rmin(rmax(v, -1.0f), 1.0f)); } // :792-794
inline float angle_between(V3 v1, V3 v2) { // :796-802
    if (dot(v1, v2) < 0.0f) return PI - 2.0f * safe_asin(magnitude(v2 + v1) / 2.0f);
    return 2.0f * safe_asin(magnitude(v2 - v1) / 2.0f);
}
inline V3 rotate_vector(float sin_theta, float cos_theta, V3 axis, V3 v) { // rotate(..).transform_vector(v), :804-824
    V3 a = normalize(axis);
    float c0r0 = a.x * a.x + (1.0f - a.x * a.x) * cos_theta;
    float c0r1 = a.x * a.y * (1.0f - cos_theta) - a.z * sin_theta;
    float c0r2 = a.x * a.z * (1.0f - cos_theta) + a.y * sin_theta;
    float c1r0 = a.x * a.y * (1.0f - cos_theta) + a.z * sin_theta;
    float c1r1 = a.y * a.y + (1.0f - a.y * a.y) * cos_theta;
    float c1r2 = a.y * a.z * (1.0f - cos_theta) - a.x * sin_theta;
    float c2r0 = a.x * a.z * (1.0f - cos_theta) - a.y * sin_theta;
    float c2r1 = a.y * a.z * (1.0f - cos_theta) + a.x * sin_theta;
    float c2r2 = a.z * a.z + (1.0f - a.z * a.z) * cos_theta;
    // Matrix4::new(..) lists columns; .transpose() makes (c0r0, c0r1, c0r2) the first ROW; M * (v, 0) = col0 v.x + col1 v.y + col2 v.z + col3 0
    return V3{c0r0 * v.x + c0r1 * v.y + c0r2 * v.z + 0.0f, c1r0 * v.x + c1r1 * v.y + c1r2 * v.z + 0.0f, c2r0 * v.x + c2r1 * v.y + c2r2 * v.z + 0.0f};
}
struct DirectionCone { // :782-900
    V3 w{0.0f, 0.0f, 1.0f};
    float cos_theta = -1.0f;
    bool empty = false;
    static DirectionCone entire_sphere() { return DirectionCone{}; }
    static DirectionCone subtended_directions(const AABB &b, V3 p) { // :840-855
        V3 c = b.center();
        float radius = magnitude(c - b.p_max); // to_sphere
        if (magnitude2(p - c) < radius * radius) return entire_sphere();
        DirectionCone r;
        r.w = normalize(c - p);
        float sin_theta_max_2 = radius * radius / magnitude2(c - p);
        r.cos_theta = std::sqrt(rmax(1.0f - sin_theta_max_2, 0.0f));
        return r;
    }
    static DirectionCone union_(const DirectionCone &a, const DirectionCone &b) { // :857-899
        if (a.empty) return b;
        if (b.empty) return a;
        float theta_a = safe_acos(a.cos_theta), theta_b = safe_acos(b.cos_theta), theta_d = angle_between(a.w, b.w);
        if (rmin(theta_d + theta_b, PI) <= theta_a) return a;
        if (rmin(theta_d + theta_a, PI) <= theta_b) return b;
        float theta_o = (theta_a + theta_d + theta_b) / 2.0f;
        if (theta_o >= PI) return entire_sphere();
        float theta_r = theta_o - theta_a;
        V3 wr = cross(a.w, b.w);
        if (magnitude2(wr) == 0.0f) return entire_sphere();
        float degrees = theta_r * 57.2957795130823208767981548141051703f; // f32::to_degrees
        float radians = degrees * (PI / 180.0f);                          // f32::to_radians
        DirectionCone r;
        r.w = rotate_vector(std::sin(radians), std::cos(radians), wr, a.w);
        r.cos_theta = std::cos(theta_o);
        return r;
    }
};
constexpr float EPSILON_ATS = 0.0001f;
struct LightBounds { // :902-1108
    AABB aabb;
    V3 w{0.0f, 0.0f, 1.0f};
    float phi = 0.0f, theta_o = 0.0f, theta_e = 0.0f, cos_theta_o = 1.0f, cos_theta_e = 1.0f;
    bool two_sided = false;
    static LightBounds union_(const LightBounds &a, const LightBounds &b) { // :948-973
        if (a.phi == 0.0f) return b;
        if (b.phi == 0.0f) return a;
        DirectionCone ca, cb;
        ca.w = a.w, ca.cos_theta = a.cos_theta_o, cb.w = b.w, cb.cos_theta = b.cos_theta_o;
        DirectionCone c = DirectionCone::union_(ca, cb);
        LightBounds r;
        r.theta_o = safe_acos(c.cos_theta);
        r.theta_e = rmax(a.theta_e, b.theta_e);
        r.aabb = a.aabb.union_aabb(b.aabb);
        r.w = c.w;
        r.phi = a.phi + b.phi;
        r.cos_theta_o = std::cos(r.theta_o);
        r.cos_theta_e = std::cos(r.theta_e);
        r.two_sided = a.two_sided | b.two_sided;
        return r;
    }
    float importance_point(V3 p, const V3 *n) const { // :1034-1108
        V3 pc = aabb.center();
        float d2 = rmax(magnitude2(p - pc), EPSILON_ATS);
        V3 wi = normalize(p - pc);
        float cos_theta = dot(w, wi);
        if (two_sided) cos_theta = std::fabs(cos_theta);
        float sin_theta = std::sqrt(rmax(1.0f - cos_theta * cos_theta, 0.0f));
        auto cos_sub_clamped = [](float sin_a, float cos_a, float sin_b, float cos_b) { return cos_a > cos_b ? 1.0f : cos_a * cos_b + sin_a * sin_b; };
        auto sin_sub_clamped = [](float sin_a, float cos_a, float sin_b, float cos_b) { return cos_a > cos_b ? 1.0f : sin_a * cos_b - cos_a * sin_b; };
        float cos_theta_u = DirectionCone::subtended_directions(aabb, p).cos_theta;
        float sin_theta_u = std::sqrt(rmax(1.0f - cos_theta_u * cos_theta_u, 0.0f));
        float sin_theta_o = std::sqrt(rmax(1.0f - cos_theta_o * cos_theta_o, 0.0f));
        float cos_theta_x = cos_sub_clamped(sin_theta, cos_theta, sin_theta_o, cos_theta_o);
        float sin_theta_x = sin_sub_clamped(sin_theta, cos_theta, sin_theta_o, cos_theta_o);
        float cos_theta_p = cos_sub_clamped(sin_theta_x, cos_theta_x, sin_theta_u, cos_theta_u);
        if (cos_theta_p <= cos_theta_e) return 0.0f;
        float imp = phi * cos_theta_p / d2;
        if (n) {
            float cos_theta_i = std::fabs(dot(wi, *n));
            float sin_theta_i = std::sqrt(rmax(1.0f - cos_theta_i * cos_theta_i, 0.0f));
            imp *= cos_sub_clamped(sin_theta_i, cos_theta_i, sin_theta_u, cos_theta_u);
        }
        return rmax(imp, 0.0f);
    }
};
struct LightProxy { // :1110-1114
    size_t emitter_id, primitive_idx;
    LightBounds bounds;
};
struct LightBVHNode { // :1116-1129
    int left = -1, right = -1, parent = -1;
    LightBounds bounds;
    int light = -1;
    bool is_leaf() const { return left < 0 && right < 0; }
};
struct LightSamplerATS { // :1130-1400
    int root = -1;
    std::vector<LightBVHNode> nodes;
    std::vector<LightProxy> lights;
    std::vector<std::pair<std::pair<size_t, size_t>, size_t>> query_to_nodes; // (emitter, primitive) -> node
    bool failed = false; // a slice came out empty: unimplemented!() in the reference
    size_t node_of(size_t emitter, size_t prim) const {
        for (auto &q : query_to_nodes)
            if (q.first.first == emitter && q.first.second == prim) return q.second;
        std::abort(); // .unwrap()
    }
    size_t build_bvh(size_t index, LightProxy *ls, size_t n) { // :1145-1287
        if (n == 0) {
            failed = true;
            return 0;
        }
        if (n == 1) {
            LightBVHNode leaf;
            leaf.bounds = ls[0].bounds, leaf.light = (int)index;
            nodes.push_back(leaf);
            query_to_nodes.push_back({{ls[0].emitter_id, ls[0].primitive_idx}, nodes.size() - 1});
            return nodes.size() - 1;
        }
        AABB bounds, centroid_bounds;
        for (size_t i = 0; i < n; i++) {
            bounds = bounds.union_aabb(ls[i].bounds.aabb);
            centroid_bounds = centroid_bounds.union_vec(ls[i].bounds.aabb.center());
        }
        auto offset = [&](V3 v, int dim) { // AABB::offset, structure.rs:809-818
            float o = comp(v - centroid_bounds.p_min, dim), sz = comp(centroid_bounds.size(), dim);
            return sz != 0.0f ? o / sz : 0.0f;
        };
        constexpr size_t NBUCKETS = 12;
        auto bucket_of = [&](const LightProxy &l, int dim) {
            size_t i = (size_t)as_usize((float)NBUCKETS * offset(l.bounds.aabb.center(), dim));
            return i < NBUCKETS - 1 ? i : NBUCKETS - 1;
        };
        float min_cost = F32_MAX;
        int min_cost_bucket = -1, min_cost_dim = -1;
        for (int dim = 0; dim < 3; dim++) {
            if (comp(centroid_bounds.p_max, dim) == comp(centroid_bounds.p_min, dim)) continue;
            std::vector<LightBounds> buckets(NBUCKETS);
            for (size_t i = 0; i < n; i++) {
                size_t b = bucket_of(ls[i], dim);
                buckets[b] = LightBounds::union_(buckets[b], ls[i].bounds);
            }
            auto momega = [](const LightBounds &b) {
                float theta_w = rmin(b.theta_o + b.theta_e, PI);
                return 2.0f * PI * (1.0f - std::cos(b.theta_o)) +
                       FRAC_PI_2 * (2.0f * theta_w * std::sin(b.theta_o) - std::cos(b.theta_o - 2.0f * theta_w) - 2.0f * b.theta_o * std::sin(b.theta_o) + std::cos(b.theta_o));
            };
            for (size_t i = 0; i + 1 < NBUCKETS; i++) {
                LightBounds b0, b1;
                for (size_t j = 0; j < i + 1; j++) b0 = LightBounds::union_(b0, buckets[j]);
                for (size_t j = i + 1; j < NBUCKETS; j++) b1 = LightBounds::union_(b1, buckets[j]);
                V3 sz = bounds.size();
                float kr = rmax(rmax(sz.x, sz.y), sz.z) / comp(sz, dim);
                float c = kr * (b0.phi * momega(b0) * b0.aabb.surface_area() + b1.phi * momega(b1) * b1.aabb.surface_area());
                if (c > 0.0f && c < min_cost) min_cost = c, min_cost_bucket = (int)i, min_cost_dim = dim;
            }
        }
        size_t mid;
        if (min_cost_dim == -1) mid = n / 2;
        else { // itertools::partition(lights.iter_mut(), pred)
            auto pred = [&](const LightProxy &l) { return bucket_of(l, min_cost_dim) <= (size_t)min_cost_bucket; };
            size_t split_index = 0, lo = 0, hi = n; // the iterator covers [lo, hi)
            while (lo < hi) {
                LightProxy &front = ls[lo++];
                if (!pred(front)) {
                    bool swapped = false;
                    while (lo < hi) {
                        LightProxy &back = ls[--hi];
                        if (pred(back)) {
                            std::swap(front, back);
                            swapped = true;
                            break;
                        }
                    }
                    if (!swapped) break; // next_back() returned None: break 'main
                }
                split_index++;
            }
            mid = split_index;
        }
        size_t left = build_bvh(index, ls, mid);
        if (failed) return 0;
        size_t right = build_bvh(index + mid, ls + mid, n - mid);
        if (failed) return 0;
        LightBVHNode inner;
        inner.left = (int)left, inner.right = (int)right;
        inner.bounds = LightBounds::union_(nodes[left].bounds, nodes[right].bounds);
        nodes.push_back(inner);
        size_t id = nodes.size() - 1;
        nodes[left].parent = (int)id, nodes[right].parent = (int)id;
        return id;
    }
    float prob_left(const LightBVHNode &node, V3 p, const V3 *n) const {
        float imp_left = nodes[node.left].bounds.importance_point(p, n), imp_right = nodes[node.right].bounds.importance_point(p, n);
        float imp_total = imp_left + imp_right;
        return (imp_left == 0.0f && imp_right == 0.0f) ? 0.5f : imp_left / imp_total;
    }
    float pdf(size_t id_emitter, size_t id_primitive, V3 p, const V3 *n) const { // :1319-1359
        size_t id = node_of(id_emitter, id_primitive);
        float pdf = 1.0f;
        while (nodes[id].parent >= 0) {
            size_t id_parent = (size_t)nodes[id].parent;
            float pl = prob_left(nodes[id_parent], p, n);
            if ((size_t)nodes[id_parent].left == id) pdf *= pl;
            else pdf *= 1.0f - pl;
            id = id_parent;
        }
        return pdf;
    }
    const LightProxy &sample(float r, V3 p, const V3 *n, float *pdf_sel_out) const { // :1361-1399
        float pdf_sel = 1.0f;
        size_t node_index = (size_t)root;
        for (;;) {
            const LightBVHNode &node = nodes[node_index];
            if (node.is_leaf()) {
                *pdf_sel_out = pdf_sel;
                return lights[node.light];
            }
            float pl = prob_left(node, p, n);
            if (r < pl) {
                r = r / pl;
                node_index = (size_t)node.left;
                pdf_sel *= pl;
            } else {
                r = (r - pl) / (1.0f - pl);
                node_index = (size_t)node.right;
                pdf_sel *= 1.0f - pl;
            }
        }
    }
};
struct EmitterSampler { // :1491-1495
    std::unique_ptr<LightSamplerATS> ats; // Some(..) after build_ats (`-x ats`)
    std::vector<std::unique_ptr<Emitter>> emitters;
    std::vector<const Mesh *> emitter_mesh; // the mesh behind emitters[i], or null
    Distribution1D emitters_cdf;
    const Emitter *of_mesh(const Mesh *m) const {
        for (size_t i = 0; i < emitters.size(); i++)
            if (emitter_mesh[i] == m) return emitters[i].get();
        return nullptr;
    }
    float pdf(const Emitter *e) const { // :1510-1526 (pointer identity)
        for (size_t i = 0; i < emitters.size(); i++)
            if (emitters[i].get() == e) return emitters_cdf.pdf(i);
        return 0.0f; // the reference panics here; unreachable (only listed emitters are queried)
    }
    // :1566-1602.  With the light tree: direct_pdf_tri * ats.pdf(emitter, primitive, importance_point(ls.o, n))
    PDF direct_pdf(const Emitter *e, const LightSamplingPDF &ls, const V3 *n = nullptr, long id_primitive = -1) const {
        if (!ats) return e->direct_pdf(ls) * pdf(e);
        size_t id_emitter = emitters.size();
        for (size_t i = 0; i < emitters.size(); i++)
            if (emitters[i].get() == e) {
                id_emitter = i;
                break;
            }
        if (id_emitter == emitters.size()) return PDF{PDF::SolidAngle, 0.0f}; // "PDF emitter without intersecting an emitter"
        if (id_primitive < 0) std::abort();                                   // id_primitive.unwrap()
        return e->direct_pdf_tri(ls, (size_t)id_primitive) * ats->pdf(id_emitter, (size_t)id_primitive, ls.o, n);
    }
    PDF direct_pdf(const Mesh *m, const LightSamplingPDF &ls, const V3 *n = nullptr, long id_primitive = -1) const { return direct_pdf(of_mesh(m), ls, n, id_primitive); }
    LightSampling sample_light(const Math &m, V3 p, const V3 *n, float r_sel, float r, P2 uv) const { // :1604-1639, :1641-1647
        if (ats) {
            float pdf_sel;
            const LightProxy &light_info = ats->sample(r_sel, p, n, &pdf_sel);
            LightSampling res = emitters[light_info.emitter_id]->direct_sample_tri(p, light_info.primitive_idx, uv);
            div_assign(res.weight, pdf_sel);
            res.pdf = res.pdf * pdf_sel;
            return res;
        }
        size_t id_light = emitters_cdf.sample_discrete(r_sel);
        float pdf_sel = emitters_cdf.pdf(id_light);
        LightSampling res = emitters[id_light]->direct_sample(m, p, r, uv);
        div_assign(res.weight, pdf_sel);
        res.pdf = res.pdf * pdf_sel;
        return res;
    }
};

// ============================================================================================
// camera.rs
// ============================================================================================
struct Camera {
    uint32_t img_x, img_y;
    M4 sample_to_camera, to_world;
    V3 position() const { return transform_point(to_world, V3{0, 0, 0}); } // :140-142
    Ray generate(P2 px) const {                                            // :81-91
        V3 near_p = transform_point(sample_to_camera, V3{px.x / (float)img_x, px.y / (float)img_y, 0.0f});
        V3 d = normalize(near_p);
        return ray_new(position(), transform_vector(to_world, d));
    }
};
M4 m4_mul(const M4 &a, const M4 &b) {
    M4 r;
    for (int c = 0; c < 4; c++) {
        float v[4] = {b.at(c, 0), b.at(c, 1), b.at(c, 2), b.at(c, 3)}, o[4];
        m4_mul_v4(a, v, o);
        for (int k = 0; k < 4; k++) r.m[4 * c + k] = o[k];
    }
    return r;
}
bool m4_invert(const M4 &a, M4 *out) { // cgmath SquareMatrix::invert: adjugate / determinant (host-only)
    auto det3 = [](float a0, float a1, float a2, float b0, float b1, float b2, float c0, float c1, float c2) {
        return a0 * (b1 * c2 - c1 * b2) - b0 * (a1 * c2 - c1 * a2) + c0 * (a1 * b2 - b1 * a2);
    };
    float cof[16];
    for (int c = 0; c < 4; c++)
        for (int r = 0; r < 4; r++) {
            float s[9];
            int k = 0;
            for (int cc = 0; cc < 4; cc++) {
                if (cc == c) continue;
                for (int rr = 0; rr < 4; rr++) {
                    if (rr == r) continue;
                    s[k++] = a.at(cc, rr);
                }
            }
            float d = det3(s[0], s[1], s[2], s[3], s[4], s[5], s[6], s[7], s[8]);
            cof[4 * c + r] = ((c + r) & 1) ? -d : d;
        }
    float det = a.at(0, 0) * cof[0] + a.at(0, 1) * cof[1] + a.at(0, 2) * cof[2] + a.at(0, 3) * cof[3];
    if (det == 0.0f) return false;
    float inv = 1.0f / det;
    for (int c = 0; c < 4; c++)
        for (int r = 0; r < 4; r++) out->m[4 * c + r] = cof[4 * r + c] * inv;
    return true;
}
bool camera_new(uint32_t w, uint32_t h, int fov_axis, float fov_deg, const M4 &to_world, bool flip, M4 *s2c) { // :31-67
    float x_v = flip ? 1.0f : -1.0f;
    float aspect_ratio = (float)w / (float)h;
    float fov_rad = fov_axis ? fov_deg * PI / 180.0f : fov_deg * aspect_ratio * PI / 180.0f;
    auto diag = [](float a, float b, float c) {
        M4 m{};
        m.m[0] = a, m.m[5] = b, m.m[10] = c, m.m[15] = 1.0f;
        return m;
    };
    M4 trans = diag(1, 1, 1);
    trans.m[12] = -1.0f, trans.m[13] = -1.0f / aspect_ratio, trans.m[14] = 0.0f;
    M4 persp{};
    float f = 1.0f / std::tan(fov_rad / 2.0f); // Rad::cot(fovy / 2)
    float near = 1e-2f, far = 1000.0f, aspect = 1.0f;
    persp.m[0] = f / aspect, persp.m[5] = f, persp.m[10] = (far + near) / (near - far), persp.m[11] = -1.0f;
    persp.m[14] = (2.0f * far * near) / (near - far);
    M4 camera_to_sample = m4_mul(m4_mul(m4_mul(diag(-0.5f, -0.5f * aspect_ratio, 1.0f), trans), persp), diag(x_v, 1.0f, -1.0f));
    (void)to_world;
    return m4_invert(camera_to_sample, s2c);
}

// ============================================================================================
// samplers: trait Sampler (samplers/mod.rs:3-9); IndependentSampler (independent.rs) over
// rand 0.8.5 SmallRng == xoshiro256++ on 64-bit targets (crate not vendored: restated from
// the published algorithm; seeding variant selectable, see oracle.h).
// ============================================================================================
inline uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
struct Xoshiro256PP {
    uint64_t s[4];
    uint64_t next_u64() {
        uint64_t result = rotl64(s[0] + s[3], 23) + s[0];
        uint64_t t = s[1] << 17;
        s[2] ^= s[0], s[3] ^= s[1], s[1] ^= s[2], s[0] ^= s[3];
        s[2] ^= t;
        s[3] = rotl64(s[3], 45);
        return result;
    }
    uint32_t next_u32() { return (uint32_t)(next_u64() >> 32); }
    // rand::distributions::Standard for f32: 24 random bits scaled by 2^-24
    float gen_f32() { return (float)(next_u32() >> 8) * (1.0f / 16777216.0f); }
    static Xoshiro256PP seed_from_u64(uint64_t state, uint32_t seeding) {
        Xoshiro256PP r;
        if (seeding == ORC_SEED_SPLITMIX64) {
            for (int i = 0; i < 4; i++) {
                state += 0x9e3779b97f4a7c15ULL;
                uint64_t z = state;
                z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
                z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
                r.s[i] = z ^ (z >> 31);
            }
        } else { // rand_core 0.6 SeedableRng::seed_from_u64: PCG32 stream fills the 32-byte seed (LE)
            uint32_t w[8];
            for (int i = 0; i < 8; i++) {
                state = state * 6364136223846793005ULL + 11634580027462260723ULL;
                uint32_t xorshifted = (uint32_t)(((state >> 18) ^ state) >> 27);
                uint32_t rot = (uint32_t)(state >> 59);
                w[i] = (xorshifted >> rot) | (xorshifted << ((32 - rot) & 31));
            }
            for (int i = 0; i < 4; i++) r.s[i] = (uint64_t)w[2 * i] | ((uint64_t)w[2 * i + 1] << 32);
        }
        if ((r.s[0] | r.s[1] | r.s[2] | r.s[3]) == 0) return seed_from_u64(0, ORC_SEED_SPLITMIX64);
        return r;
    }
};
struct Sampler {
    virtual ~Sampler() = default;
    virtual float next() = 0;
    P2 next2d() { // x then y, independent.rs:13-17
        float x = next();
        float y = next();
        return P2{x, y};
    }
    uint32_t draws = 0;
};
struct IndependentSampler : Sampler { // mode A
    Xoshiro256PP rnd;
    uint32_t seeding;
    float next() override {
        draws++;
        return rnd.gen_f32();
    }
    IndependentSampler clone_box() { // independent.rs:18-22
        IndependentSampler c;
        c.seeding = seeding;
        c.rnd = Xoshiro256PP::seed_from_u64(rnd.next_u64(), seeding);
        return c;
    }
};
// mode B: counter-based stream per (seed, pixel, sample); DESIGN.md §rng.  Same on the GPU.
inline uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
struct CounterSampler : Sampler {
    uint64_t key;
    uint32_t n = 0;
    CounterSampler(uint64_t seed, uint32_t pixel, uint32_t sample) {
        uint64_t h = mix64(seed + 0x9e3779b97f4a7c15ULL);
        key = mix64(h ^ (((uint64_t)pixel << 32) | (uint64_t)sample));
    }
    float next() override {
        // draw n (1-based) = one 24-bit half of hash (n + 1) / 2: bits 40..63 for odd n, bits 16..39 for even n
        draws++;
        n++;
        uint64_t z = mix64(key + (uint64_t)((n + 1u) >> 1) * 0x9e3779b97f4a7c15ULL);
        uint32_t bits = (n & 1u) ? (uint32_t)(z >> 40) : ((uint32_t)(z >> 16) & 0xffffffu);
        return (float)bits * (1.0f / 16777216.0f);
    }
};

// ============================================================================================
// Scene (scene.rs) + accel.rs
// ============================================================================================
struct Counters {
    uint64_t segments = 0, shadow_rays = 0, shadow_visible = 0, hits = 0, max_depth = 0, nee_added = 0;
};
struct TriRef {
    uint32_t id_mesh, id_tri;
};
struct BVHNode { // accel.rs:79-88
    AABB aabb;
    size_t info, count;
    bool is_leaf() const { return count != 0; }
};
struct CachedAABB {
    AABB aabb;
    TriRef info;
};
struct Scene {
    Camera camera;
    std::vector<std::unique_ptr<Mesh>> meshes;
    EmitterSampler emitters;
    // BVHAccel
    std::vector<TriRef> primitives;
    std::vector<BVHNode> nodes;

    std::vector<rl_light_desc> lights; // Scene.emitters: EmittersState::Unbuild (point / directional), in file order
    bool has_environment = false;      // Scene.emitter_environment: EnvironmentLight with a constant colour
    Color environment{};
    const EnvironmentLight *env_emitter = nullptr;
    BitmapTex environment_image; // EnvironmentLightColor::Texture when size_x != 0
    Color enviroment_luminance(const Math &m, V3 d) const { return env_emitter ? env_emitter->luminance.eval(m, d) : Color::zero(); } // scene.rs:125-130
    BoundingSphere bsphere{};
    bool build_ats = false; // Scene::build_emitters(build_ats), scene.rs:53, 118-120
    std::string ats_error;
    void build_emitters() { // scene.rs:53-123
        // bounding sphere: union of Mesh::compute_aabb (all vertices, geometry.rs:441-456) and the camera position
        AABB aabb;
        for (auto &m : meshes) {
            AABB a;
            for (auto &v : m->vertices) a = a.union_vec(v);
            V3 sz = a.size();
            if (sz.x < EPSILON) a.p_max.x += EPSILON, a.p_min.x -= EPSILON;
            if (sz.y < EPSILON) a.p_max.y += EPSILON, a.p_min.y -= EPSILON;
            if (sz.z < EPSILON) a.p_max.z += EPSILON, a.p_min.z -= EPSILON;
            aabb = aabb.union_aabb(a);
        }
        aabb = aabb.union_vec(camera.position());
        V3 c = aabb.center();
        bsphere = BoundingSphere{c, magnitude(c - aabb.p_max)}; // AABB::to_sphere, structure.rs:871-877
        emitters.emitters.clear();
        emitters.emitter_mesh.clear();
        for (auto &m : meshes)
            if (m->is_light()) {
                emitters.emitters.push_back(std::make_unique<MeshEmitter>(m.get()));
                emitters.emitter_mesh.push_back(m.get());
            }
        env_emitter = nullptr;
        if (has_environment) { // scene.rs:69-81: preprocess, then pushed right after the mesh lights
            auto e = std::make_unique<EnvironmentLight>();
            if (environment_image.size_x != 0) e->luminance.new_texture(environment_image);
            else e->luminance.constant = environment;
            e->bsphere = bsphere;
            e->bsphere.radius *= 1.1f;
            env_emitter = e.get();
            emitters.emitters.push_back(std::move(e));
            emitters.emitter_mesh.push_back(nullptr);
        }
        for (auto &l : lights) { // e.preprocess(self); emitters.push(e)  (scene.rs:85-96)
            if (l.kind == RL_LIGHT_POINT) {
                auto e = std::make_unique<PointEmitter>();
                e->intensity = Color{l.intensity[0], l.intensity[1], l.intensity[2]}, e->position = V3{l.v[0], l.v[1], l.v[2]};
                emitters.emitters.push_back(std::move(e));
            } else {
                auto e = std::make_unique<DirectionalLight>();
                e->intensity = Color{l.intensity[0], l.intensity[1], l.intensity[2]}, e->direction = V3{l.v[0], l.v[1], l.v[2]};
                e->bsphere = bsphere;
                e->bsphere.radius *= 1.1f;
                emitters.emitters.push_back(std::move(e));
            }
            emitters.emitter_mesh.push_back(nullptr);
        }
        if (emitters.emitters.empty()) return;
        std::vector<float> fl;
        for (auto &e : emitters.emitters) fl.push_back(e->flux().channel_max());
        emitters.emitters_cdf = Distribution1D::normalize(fl);
        emitters.ats.reset();
        if (build_ats) { // EmitterSampler::build_ats -> LightSamplerATS::new (emitter.rs:1290-1317, 1505-1508)
            auto ats = std::make_unique<LightSamplerATS>();
            for (size_t i = 0; i < emitters.emitters.size(); i++) {
                if (!emitters.emitters[i]->is_surface()) { // assert!(e.is_surface())
                    ats_error = "ats: every emitter must be a surface (assert!(e.is_surface()), emitter.rs:1292-1294)";
                    return;
                }
                const Mesh *m = emitters.emitter_mesh[i];
                for (size_t t = 0; t < m->indices.size(); t++) { // Mesh::convert_light_proxy, emitter.rs:730-779
                    Idx3 idx = m->indices[t];
                    V3 v0 = m->vertices[idx.x], v1 = m->vertices[idx.y], v2 = m->vertices[idx.z];
                    V3 n = cross(v1 - v0, v2 - v0);
                    LightProxy lp;
                    lp.emitter_id = i, lp.primitive_idx = t;
                    lp.bounds.w = normalize(n);
                    lp.bounds.theta_o = 0.0f, lp.bounds.theta_e = FRAC_PI_2;
                    UV cuv; // "For now we interpolate at the middle" (:742-750)
                    if (m->has_uv) {
                        P2 a = m->uv[idx.x], b = m->uv[idx.y], c = m->uv[idx.z];
                        cuv.some = true, cuv.v = P2{((a.x + b.x) + c.x) / 3.0f, ((a.y + b.y) + c.y) / 3.0f};
                    }
                    lp.bounds.phi = m->emit(cuv).channel_max() * magnitude(n) * 0.5f;
                    lp.bounds.aabb = AABB{}.union_vec(v0).union_vec(v1).union_vec(v2);
                    lp.bounds.cos_theta_o = std::cos(lp.bounds.theta_o), lp.bounds.cos_theta_e = std::cos(lp.bounds.theta_e);
                    ats->lights.push_back(lp);
                }
            }
            if (ats->lights.empty()) {
                ats_error = "ats: no emissive triangle (LightSamplerATS::new(..).unwrap())";
                return;
            }
            ats->root = (int)ats->build_bvh(0, ats->lights.data(), ats->lights.size());
            if (ats->failed) {
                ats_error = "ats: a split left one side empty (unimplemented!() in build_bvh)";
                return;
            }
            emitters.ats = std::move(ats);
        }
    }

    // ---- BVHAccel::new, accel.rs:201-240 -----------------------------------------------
    static AABB compute_aabb(const std::vector<CachedAABB> &aabbs, size_t start, size_t count) { // :107-113
        AABB a;
        for (size_t i = 0; i < count; i++) a = a.union_aabb(aabbs[i + start].aabb);
        return a;
    }
    void subdivide_node(size_t id_node, std::vector<CachedAABB> &aabbs) { // :115-199
        if (nodes[id_node].count <= 2) return;
        size_t nb_prim = nodes[id_node].count, first_prim = nodes[id_node].info;
        nodes[id_node].count = 0;
        nodes[id_node].info = nodes.size();
        size_t best_pos = 0;
        float best_cost = std::numeric_limits<float>::infinity();
        int best_axis = 3;
        auto by_axis = [&](int o) {
            // Rust's sort_by is a stable merge sort that only asks `compare(a,b) == Less`; the
            // reference comparator (never Equal) therefore acts as a stable sort on `<`.
            std::stable_sort(aabbs.begin() + first_prim, aabbs.begin() + first_prim + nb_prim,
                             [o](const CachedAABB &a, const CachedAABB &b) { return comp(a.aabb.center(), o) < comp(b.aabb.center(), o); });
        };
        {
            std::vector<float> scores(nb_prim - 1, 0.0f);
            for (int o = 0; o < 3; o++) {
                by_axis(o);
                AABB tmp;
                for (size_t id = 0; id < nb_prim - 1; id++) {
                    size_t id_left = nb_prim - id - 1;
                    tmp = tmp.union_aabb(aabbs[id_left + first_prim].aabb);
                    scores[id_left - 1] = tmp.surface_area() * (float)(id + 1);
                }
                tmp = AABB();
                for (size_t id = 0; id < nb_prim - 1; id++) {
                    tmp = tmp.union_aabb(aabbs[id + first_prim].aabb);
                    scores[id] += tmp.surface_area() * (float)(id + 1);
                    if (scores[id] < best_cost) {
                        best_cost = scores[id];
                        best_axis = o;
                        best_pos = id + 1;
                    }
                }
            }
        }
        if (best_axis < 3) by_axis(best_axis);
        size_t offset;
        if (best_pos == nb_prim || best_pos == 0) offset = std::max<size_t>((size_t)((float)nb_prim * 0.5f), 1);
        else offset = best_pos;
        BVHNode left{compute_aabb(aabbs, first_prim, offset), first_prim, offset};
        BVHNode right{compute_aabb(aabbs, first_prim + offset, nb_prim - offset), first_prim + offset, nb_prim - offset};
        size_t id_left = nodes.size();
        nodes.push_back(left);
        nodes.push_back(right);
        subdivide_node(id_left, aabbs);
        subdivide_node(id_left + 1, aabbs);
    }
    void build_bvh() {
        AABB root_aabb;
        std::vector<CachedAABB> cached;
        for (size_t m = 0; m < meshes.size(); m++)
            for (size_t i = 0; i < meshes[m]->indices.size(); i++) {
                cached.push_back(CachedAABB{meshes[m]->compute_aabb_tri(i), TriRef{(uint32_t)m, (uint32_t)i}});
                root_aabb = root_aabb.union_aabb(cached.back().aabb);
            }
        nodes.clear();
        nodes.push_back(BVHNode{root_aabb, 0, cached.size()});
        subdivide_node(0, cached);
        primitives.clear();
        for (auto &c : cached) primitives.push_back(c.info);
    }

    // ---- BVHAccel::intersect, accel.rs:243-288 ------------------------------------------
    bool bvh_intersect(size_t id_node, const Ray &ray, IntersectionUV &its, TriRef *res) const {
        const BVHNode &node = nodes[id_node];
        if (node.is_leaf()) {
            bool found = false;
            for (size_t k = 0; k < node.count; k++) {
                const TriRef &e = primitives[node.info + k];
                if (meshes[e.id_mesh]->intersection_tri(e.id_tri, ray.o, ray.d, its)) {
                    *res = e;
                    found = true;
                }
            }
            return found;
        }
        size_t id1 = node.info, id2 = node.info + 1;
        float d1, d2;
        if (!nodes[id1].aabb.intersect(ray, &d1)) d1 = std::numeric_limits<float>::infinity();
        if (!nodes[id2].aabb.intersect(ray, &d2)) d2 = std::numeric_limits<float>::infinity();
        if (d1 > d2) {
            std::swap(d1, d2);
            std::swap(id1, id2);
        }
        bool found = false;
        if (d1 < its.t) found = bvh_intersect(id1, ray, its, res);
        if (d2 < its.t) {
            TriRef r2;
            if (bvh_intersect(id2, ray, its, &r2)) {
                *res = r2;
                found = true;
            }
        }
        return found;
    }

    // ---- Acceleration::trace ---------------------------------------------------------------
    // BVH: accel.rs:292-315.  NAIVE: the brute-force loop of accel.rs:23-51 (mesh-major, strict
    // `t < its.t`, so the lowest (mesh,tri) wins exact ties) behind BVHAccel's root-box test.
    bool trace_uv(const Ray &ray, uint32_t accel_mode, IntersectionUV &its, TriRef *res) const {
        float t0;
        if (!nodes[0].aabb.intersect(ray, &t0)) return false;
        its = IntersectionUV{F32_MAX, V3{0, 0, 0}, V3{0, 0, 0}, 0.0f, 0.0f};
        if (accel_mode == ORC_ACCEL_NAIVE) {
            for (size_t m = 0; m < meshes.size(); m++)
                for (size_t i = 0; i < meshes[m]->indices.size(); i++)
                    if (meshes[m]->intersection_tri(i, ray.o, ray.d, its)) *res = TriRef{(uint32_t)m, (uint32_t)i};
        } else {
            bvh_intersect(0, ray, its, res);
        }
        return its.t != F32_MAX;
    }
    bool trace(const Ray &ray, uint32_t accel_mode, Counters &c, Intersection *out) const {
        c.segments++;
        IntersectionUV its;
        TriRef res{0, 0};
        if (!trace_uv(ray, accel_mode, its, &res)) return false;
        c.hits++;
        *out = fill_intersection(meshes[res.id_mesh].get(), res.id_tri, its.u, its.v, ray, its.n, its.t, its.p);
        return true;
    }
    // ---- Acceleration::visible, accel.rs:316-343 (NAIVE: loop of :52-76 behind the root test) --
    bool visible(V3 p0, V3 p1, uint32_t accel_mode, Counters &c) const {
        c.shadow_rays++;
        const float SHADOW_EPS = 0.00001f;
        V3 d = p1 - p0;
        float length = magnitude(d);
        d = d / length;
        IntersectionUV its{length * (1.0f - SHADOW_EPS), V3{0, 0, 0}, V3{0, 0, 0}, 0.0f, 0.0f};
        Ray ray{p0, d, EPSILON, length * (1.0f - SHADOW_EPS)};
        float t0;
        if (!nodes[0].aabb.intersect(ray, &t0)) return false;
        bool vis;
        if (accel_mode == ORC_ACCEL_NAIVE) {
            vis = true;
            for (size_t m = 0; m < meshes.size() && vis; m++)
                for (size_t i = 0; i < meshes[m]->indices.size(); i++)
                    if (meshes[m]->intersection_tri(i, p0, d, its)) {
                        vis = false;
                        break;
                    }
        } else {
            TriRef r;
            vis = !bvh_intersect(0, ray, its, &r);
        }
        if (vis) c.shadow_visible++;
        return vis;
    }
};

// ============================================================================================
// paths/: Vertex (vertex.rs:9-39), Edge (edge.rs:11-24), Path (path.rs:15-18)
// ============================================================================================
struct Ctx { // what the reference threads through as (accel, scene, sampler)
    const Scene *scene;
    uint32_t accel_mode;
    Math math;
    Counters *counters;
};
struct Vertex {
    enum Kind { Sensor, Surface, Light } kind;
    // Sensor
    P2 uv{};
    V3 pos{};
    // Surface
    Intersection its{};
    // Light
    V3 n{};
    UV light_uv{};
    const Emitter *emitter = nullptr;
    // edges
    int edge_in = -1;
    std::vector<int> edge_out; // Sensor/Light hold at most one
    V3 position() const { return kind == Surface ? its.p : pos; }                                   // vertex.rs:46-53
    bool on_light_source() const { return kind == Surface ? its.mesh->is_light() : kind == Light; } // :60-66
};
struct Edge {
    bool has_dist = false;
    float dist = 0;
    V3 d{};
    int v0 = -1, v1 = -1; // vertices: (VertexID, Option<VertexID>)
    PDF pdf_direction{PDF::SolidAngle, 0};
    Color weight{};
    bool has_contrib = false;
    Color contrib{};
    float rr_weight = 1;
    size_t id_sampling = 0;
};
struct Path {
    std::vector<Vertex> vertices;
    std::vector<Edge> edges;
    int register_vertex(Vertex v) {
        vertices.push_back(std::move(v));
        return (int)vertices.size() - 1;
    }
    int register_edge(Edge e) {
        edges.push_back(e);
        return (int)edges.size() - 1;
    }
};
// Vertex::contribution, vertex.rs:69-82
Color vertex_contribution(const Vertex &v, const Edge &edge) {
    if (v.kind == Vertex::Surface) {
        if (dot(v.its.n_s, -edge.d) >= 0.0f) return v.its.mesh->emit(v.its.uv);
        return Color::zero();
    }
    if (v.kind == Vertex::Light) return v.emitter->eval(v.light_uv); // emitter.eval(-edge.d, uv), emitter.rs:605-607
    return Color::zero();
}
// Edge::contribution, edge.rs:201-210 (environment luminance is zero on this path)
Color edge_contribution(const Edge &e, const Path &path, const Scene *scene, const Math &math) {
    if (e.v1 >= 0) {
        if (e.has_contrib) return e.contrib * e.weight * e.rr_weight;
        return e.weight * e.rr_weight * vertex_contribution(path.vertices[e.v1], e);
    }
    return e.weight * e.rr_weight * scene->enviroment_luminance(math, e.d); // scene.enviroment_luminance(self.d)
}
bool edge_next_on_light_source(const Edge &e, const Path &path, const Scene *scene) { // edge.rs:191-197
    if (e.v1 >= 0) return path.vertices[e.v1].on_light_source();
    return scene->has_environment;
}
// Edge::from_ray, edge.rs:65-189 (medium == None)
std::pair<int, int> edge_from_ray(Path &path, const Ray &ray, int org, PDF pdf_direction, Color weight, float rr_weight, const Ctx &cx, size_t id_sampling) {
    Edge e;
    e.d = ray.d, e.v0 = org, e.pdf_direction = pdf_direction, e.weight = weight, e.rr_weight = rr_weight, e.id_sampling = id_sampling;
    int eid = path.register_edge(e);
    Intersection its;
    if (!cx.scene->trace(ray, cx.accel_mode, *cx.counters, &its)) return {eid, -1};
    Vertex nv;
    nv.kind = Vertex::Surface;
    nv.its = its;
    nv.edge_in = eid;
    int vid = path.register_vertex(nv);
    path.edges[eid].has_dist = true;
    path.edges[eid].dist = its.dist;
    path.edges[eid].v1 = vid;
    return {eid, vid};
}
// Edge::from_vertex, edge.rs:27-63
int edge_from_vertex(Path &path, int org, PDF pdf_direction, Color weight, Color contrib, float rr_weight, int next, size_t id_sampling) {
    V3 d = path.vertices[next].position() - path.vertices[org].position();
    float dist = magnitude(d);
    d = d / dist;
    Edge e;
    e.has_dist = true, e.dist = dist, e.d = d, e.v0 = org, e.v1 = next, e.pdf_direction = pdf_direction, e.weight = weight;
    e.has_contrib = true, e.contrib = contrib, e.rr_weight = rr_weight, e.id_sampling = id_sampling;
    int eid = path.register_edge(e);
    path.vertices[next].edge_in = eid;
    return eid;
}

// ============================================================================================
// paths/strategies
// ============================================================================================
struct SamplingStrategy { // strategies/mod.rs:11-33
    virtual ~SamplingStrategy() = default;
    // returns true and fills (new_vertex, new_throughput) for Some(..)
    virtual bool sample(Path &path, int vertex_id, const Ctx &cx, Color throughput, Sampler &sampler, size_t id_strategy, uint32_t depth, int *new_vertex, Color *new_throughput) const = 0;
    virtual bool pdf(const Path &path, const Ctx &cx, int vertex_id, int edge_id, float *out) const = 0;
};
struct DirectionalSamplingStrategy : SamplingStrategy { // strategies/directional.rs
    int32_t rr_depth; // Option<u32>, -1 = None
    // bounce, :13-210 (Sensor and Surface arms; transport == Importance)
    void bounce(Path &path, int vertex_id, const Ctx &cx, Color &throughput, Sampler &sampler, size_t id_strategy, uint32_t depth, int *edge, int *new_vertex) const {
        *edge = -1, *new_vertex = -1;
        const Vertex &v = path.vertices[vertex_id];
        if (v.kind == Vertex::Sensor) { // :26-43
            Ray ray = cx.scene->camera.generate(v.uv);
            auto r = edge_from_ray(path, ray, vertex_id, PDF{PDF::SolidAngle, 1.0f}, Color::one(), 1.0f, cx, id_strategy);
            *edge = r.first, *new_vertex = r.second;
            return;
        }
        if (v.kind == Vertex::Surface) { // :44-107
            Intersection its = v.its; // copy: `path` is mutated below
            SampledDirection sb;
            P2 s2 = sampler.next2d();
            if (!its.mesh->bsdf->sample(cx.math, its.uv, its.wi, s2, &sb)) return;
            V3 d_out_global = its.frame.to_world(sb.d);
            throughput = throughput * sb.weight; // *throughput *= &weight
            if (throughput.is_zero()) return;
            bool do_rr = rr_depth < 0 ? true : (uint32_t)rr_depth <= depth;
            float rr_weight;
            if (do_rr) {
                float q = rmin(throughput.channel_max(), 0.95f);
                if (q < sampler.next()) return;
                rr_weight = 1.0f / q;
            } else rr_weight = 1.0f;
            throughput.scale(rr_weight);
            Ray ray = spawn_ray(its, d_out_global);
            auto r = edge_from_ray(path, ray, vertex_id, sb.pdf, sb.weight, rr_weight, cx, id_strategy);
            *edge = r.first, *new_vertex = r.second;
            return;
        }
        // Vertex::Light arm (light tracing) is not reachable from the sensor
    }
    bool sample(Path &path, int vertex_id, const Ctx &cx, Color throughput, Sampler &sampler, size_t id_strategy, uint32_t depth, int *new_vertex, Color *new_throughput) const override { // :211-257
        int edge, nv;
        bounce(path, vertex_id, cx, throughput, sampler, id_strategy, depth, &edge, &nv);
        if (edge >= 0) {
            Vertex &v = path.vertices[vertex_id];
            if (v.kind == Vertex::Sensor || v.kind == Vertex::Light) {
                v.edge_out.clear();
                v.edge_out.push_back(edge);
            } else v.edge_out.push_back(edge);
        }
        if (nv >= 0) {
            *new_vertex = nv;
            *new_throughput = throughput;
            return true;
        }
        return false;
    }
    bool pdf(const Path &path, const Ctx &cx, int vertex_id, int edge_id, float *out) const override { // :258-304
        const Edge &edge = path.edges[edge_id];
        if (!edge_next_on_light_source(edge, path, cx.scene)) return false;
        const Vertex &v = path.vertices[vertex_id];
        if (v.kind == Vertex::Surface) {
            if (v.its.mesh->bsdf->is_smooth()) return false;
            *out = v.its.mesh->bsdf->pdf(cx.math, v.its.uv, v.its.wi, v.its.frame.to_local(edge.d)).value();
            return true;
        }
        if (v.kind == Vertex::Sensor) {
            *out = 1.0f;
            return true;
        }
        return false;
    }
};
struct LightSamplingStrategy : SamplingStrategy { // strategies/emitters.rs
    bool sample(Path &path, int vertex_id, const Ctx &cx, Color, Sampler &sampler, size_t id_strategy, uint32_t, int *, Color *) const override { // :95-248
        if (path.vertices[vertex_id].kind != Vertex::Surface) return false;
        Intersection its = path.vertices[vertex_id].its;
        if (its.mesh->bsdf->is_smooth()) return false;
        float r_sel = sampler.next();
        float r = sampler.next();
        P2 uv = sampler.next2d();
        LightSampling rec = cx.scene->emitters.sample_light(cx.math, its.p, &its.n_s, r_sel, r, uv);
        bool visible = cx.scene->visible(its.p, rec.p, cx.accel_mode, *cx.counters); // evaluated before the && (:125-126)
        if (rec.is_valid() && visible) {
            Vertex nv;
            nv.kind = Vertex::Light;
            nv.pos = rec.p, nv.n = rec.n, nv.emitter = rec.emitter, nv.light_uv = rec.uv;
            Color weight = its.mesh->bsdf->eval(cx.math, its.uv, its.wi, its.frame.to_local(rec.d));
            int nvid = path.register_vertex(nv);
            int eid = edge_from_vertex(path, vertex_id, rec.pdf, weight, rec.weight, 1.0f, nvid, id_strategy);
            path.vertices[vertex_id].edge_out.push_back(eid);
        }
        return false; // "Finish the sampling here"
    }
    bool pdf_emitter(const Path &path, const Ctx &cx, const Ray &ray, int next_vertex_id, float *out) const { // :10-92
        if (next_vertex_id < 0) { // :18-46: the edge left the scene
            const EnvironmentLight *env = cx.scene->env_emitter;
            if (!env) return false;
            float t;
            if (!bsphere_intersect(env->bsphere, ray, &t)) std::abort(); // t.unwrap()
            V3 p = ray.o + ray.d * t;
            V3 n = normalize(env->bsphere.center - p);
            PDF pdf = cx.scene->emitters.direct_pdf(env, LightSamplingPDF{ray.o, p, n, ray.d, cx.math.mode});
            *out = pdf.value();
            return true;
        }
        const Vertex &nv = path.vertices[next_vertex_id];
        if (nv.kind == Vertex::Surface) {
            PDF p = cx.scene->emitters.direct_pdf(nv.its.mesh, LightSamplingPDF{ray.o, nv.its.p, nv.its.n_g, ray.d}, nullptr, (long)nv.its.primitive_id);
            *out = p.value();
            return true;
        }
        if (nv.kind == Vertex::Light) {
            PDF p = cx.scene->emitters.direct_pdf(nv.emitter, LightSamplingPDF{ray.o, nv.pos, nv.n, ray.d});
            *out = p.value();
            return true;
        }
        return false;
    }
    bool pdf(const Path &path, const Ctx &cx, int vertex_id, int edge_id, float *out) const override { // :250-282
        const Edge &edge = path.edges[edge_id];
        if (!edge_next_on_light_source(edge, path, cx.scene)) return false;
        const Vertex &v = path.vertices[vertex_id];
        if (v.kind == Vertex::Surface) {
            if (v.its.mesh->bsdf->is_smooth()) return false;
            Ray ray = ray_new(v.position(), edge.d);
            return pdf_emitter(path, cx, ray, edge.v1, out);
        }
        return false;
    }
};

// generate, strategies/mod.rs:35-80
struct TechniquePathTracing { // path.rs:22-35
    int32_t max_depth; // Option<u32>
    std::vector<const SamplingStrategy *> samplings;
    bool single_scattering;
    bool expand(uint32_t depth) const { return max_depth < 0 ? true : depth < (uint32_t)max_depth; }
};
void generate(Path &path, int root, const Ctx &cx, Sampler &sampler, const TechniquePathTracing &technique) {
    std::vector<std::pair<int, Color>> curr{{root, Color::one()}}, next;
    uint32_t depth = 1;
    while (!curr.empty()) {
        next.clear();
        for (auto &cv : curr) {
            if (technique.expand(depth)) {
                for (size_t id_sampling = 0; id_sampling < technique.samplings.size(); id_sampling++) {
                    int nv;
                    Color nt;
                    if (technique.samplings[id_sampling]->sample(path, cv.first, cx, cv.second, sampler, id_sampling, depth, &nv, &nt)) next.push_back({nv, nt});
                }
            }
        }
        std::swap(curr, next);
        if (depth > cx.counters->max_depth) cx.counters->max_depth = depth;
        depth++;
    }
}

// TechniquePathTracing::evalute_edge / evaluate, path.rs:37-185
Color evalute_edge(const TechniquePathTracing &tq, uint32_t curr_depth, int32_t min_depth, const Path &path, const Ctx &cx, int vertex_id, int edge_id, uint32_t strategy) {
    const Edge &edge = path.edges[edge_id];
    Color contrib = edge_contribution(edge, path, cx.scene, cx.math);
    if (strategy == RL_STRATEGY_BSDF && edge.id_sampling != 0) contrib = Color::zero();
    if (strategy == RL_STRATEGY_EMITTER && edge.id_sampling != 1) contrib = Color::zero();
    bool add_contrib = min_depth < 0 ? true : curr_depth >= (uint32_t)min_depth;
    if (!contrib.is_zero() && add_contrib) {
        float weight = 1.0f;
        if (strategy == RL_STRATEGY_ALL) {
            if (edge.pdf_direction.kind == PDF::SolidAngle) { // balance heuristic
                float v = edge.pdf_direction.v;
                float total = 0.0f;
                for (size_t id = 0; id < tq.samplings.size(); id++) {
                    float pdf;
                    if (id == edge.id_sampling) pdf = v;
                    else if (!tq.samplings[id]->pdf(path, cx, vertex_id, edge_id, &pdf)) pdf = 0.0f;
                    total += pdf;
                }
                weight = v / total;
            }
        }
        return contrib * weight;
    }
    return Color::zero();
}
Color evaluate(const TechniquePathTracing &tq, uint32_t curr_depth, int32_t min_depth, const Path &path, const Ctx &cx, int vertex_id, uint32_t strategy) {
    const Vertex &v = path.vertices[vertex_id];
    if (tq.single_scattering && (v.kind == Vertex::Surface || v.kind == Vertex::Light)) return Color::zero();
    Color l_i = Color::zero();
    if (v.kind == Vertex::Surface) {
        for (int edge_id : v.edge_out) {
            l_i = l_i + evalute_edge(tq, curr_depth, min_depth, path, cx, vertex_id, edge_id, strategy);
            const Edge &edge = path.edges[edge_id];
            if (edge.v1 >= 0) l_i = l_i + edge.weight * edge.rr_weight * evaluate(tq, curr_depth + 1, min_depth, path, cx, edge.v1, strategy);
        }
    } else if (v.kind == Vertex::Sensor) {
        const Edge &edge = path.edges[v.edge_out.at(0)];
        bool add_contrib = min_depth < 0 ? true : curr_depth >= (uint32_t)min_depth;
        Color contrib = edge_contribution(edge, path, cx.scene, cx.math);
        if (!contrib.is_zero() && add_contrib) l_i = l_i + contrib;
        if (edge.v1 >= 0) l_i = l_i + edge.weight * edge.rr_weight * evaluate(tq, curr_depth + 1, min_depth, path, cx, edge.v1, strategy);
    }
    return l_i;
}

// ============================================================================================
// integrators
// ============================================================================================
float mis_weight(float pdf_a, float pdf_b) { // integrators/mod.rs:462-478
    if (pdf_a == 0.0f) return 0.0f;
    if (!std::isfinite(pdf_a) || !std::isfinite(pdf_b)) return 0.0f;
    float w = (pdf_a * pdf_a) / ((pdf_a * pdf_a) + (pdf_b * pdf_b)); // powi(2) == x*x
    return std::isfinite(w) ? w : 0.0f;
}

// IntegratorPathTracing::compute_pixel, explicit/path.rs:197-238
Color path_compute_pixel(const rl_integrator_desc &I, uint32_t ix, uint32_t iy, const Ctx &cx, Sampler &sampler) {
    DirectionalSamplingStrategy dir;
    dir.rr_depth = I.rr_depth;
    LightSamplingStrategy light;
    TechniquePathTracing technique;
    technique.max_depth = I.max_depth;
    technique.single_scattering = I.single_scattering != 0;
    technique.samplings.push_back(&dir);
    if (I.strategy == RL_STRATEGY_ALL || I.strategy == RL_STRATEGY_EMITTER) technique.samplings.push_back(&light);
    Path path;
    Vertex root; // Path::from_sensor, paths/path.rs:56-73
    root.kind = Vertex::Sensor;
    float jx = sampler.next();
    float jy = sampler.next();
    root.uv = P2{(float)ix + jx, (float)iy + jy};
    root.pos = cx.scene->camera.position();
    int rid = path.register_vertex(root);
    generate(path, rid, cx, sampler, technique);
    return evaluate(technique, 0, I.min_depth, path, cx, rid, I.strategy);
}

// The same estimator written as a forward accumulation (throughput carried along the path), in
// the operation order of the GPU kernels.  Identical random numbers,
// rays and discrete decisions as the graph version; radiance differs only by f32 rounding of
// the re-associated products.
Color path_compute_pixel_stream(const rl_integrator_desc &I, uint32_t ix, uint32_t iy, const Ctx &cx, Sampler &sampler) {
    const Scene &sc = *cx.scene;
    const bool use_nee = (I.strategy == RL_STRATEGY_ALL || I.strategy == RL_STRATEGY_EMITTER);
    const bool mute = I.single_scattering != 0; // evaluate() returns zero for every surface vertex (path.rs:122-124)
    auto expand = [&](uint32_t depth) { return I.max_depth < 0 ? true : depth < (uint32_t)I.max_depth; };
    auto add_ok = [&](uint32_t curr_depth) { return I.min_depth < 0 ? true : curr_depth >= (uint32_t)I.min_depth; };
    Color L = Color::zero();
    float jx = sampler.next();
    float jy = sampler.next();
    uint32_t depth = 1; // generate() depth at which the current ray was sampled
    if (depth > cx.counters->max_depth) cx.counters->max_depth = depth;
    if (!expand(depth)) return L;
    Ray ray = sc.camera.generate(P2{(float)ix + jx, (float)iy + jy});
    Color T = Color::one(); // throughput carried by `ray`
    float pdf_prev = 1.0f;  // solid-angle pdf of the direction of `ray`
    bool mis_prev = true;   // false: the edge is PDF::Discrete, or was sampled at a smooth vertex (the light strategy's pdf is None: v / (v + 0))
    for (;;) {
        Intersection its;
        if (!sc.trace(ray, cx.accel_mode, *cx.counters, &its)) {
            // edge without a next vertex: weight * rr * environment luminance (edge.rs:208), same gates and MIS as an emitter hit
            if (sc.has_environment) {
                const Color env_l = sc.enviroment_luminance(cx.math, ray.d);
                if (depth == 1) {
                    if (add_ok(0) && !env_l.is_zero()) L = L + env_l;
                } else if (!mute && add_ok(depth - 1) && I.strategy != RL_STRATEGY_EMITTER) {
                    Color contrib = T * env_l;
                    if (!contrib.is_zero()) {
                        float w = 1.0f;
                        if (I.strategy == RL_STRATEGY_ALL && mis_prev) { // pdf_emitter's environment arm (emitters.rs:18-46)
                            float pl = sc.emitters.direct_pdf(sc.env_emitter, LightSamplingPDF{ray.o, V3{0, 0, 0}, V3{0, 0, 0}, ray.d, cx.math.mode}).value();
                            w = pdf_prev / (pdf_prev + pl);
                        }
                        L = L + contrib * w;
                    }
                }
            }
            break;
        }
        const BSDF &bsdf = *its.mesh->bsdf;
        // ---- emission carried by the arriving edge ----
        if (depth == 1) { // sensor edge: un-weighted (path.rs:152-165)
            if (add_ok(0) && dot(its.n_s, -ray.d) >= 0.0f && its.mesh->is_light() && !its.mesh->emit(its.uv).is_zero()) L = L + its.mesh->emit(its.uv);
        } else if (!mute && add_ok(depth - 1) && I.strategy != RL_STRATEGY_EMITTER) {
            if (dot(its.n_s, -ray.d) >= 0.0f && its.mesh->is_light()) {
                Color contrib = T * its.mesh->emit(its.uv);
                if (!contrib.is_zero()) {
                    float w = 1.0f;
                    if (I.strategy == RL_STRATEGY_ALL && mis_prev) { // balance heuristic (path.rs:78-99)
                        float pl = sc.emitters.direct_pdf(its.mesh, LightSamplingPDF{ray.o, its.p, its.n_g, ray.d}, nullptr, (long)its.primitive_id).value();
                        w = pdf_prev / (pdf_prev + pl);
                    }
                    L = L + contrib * w;
                }
            }
        }
        // ---- expand this vertex ----
        uint32_t vdepth = depth; // == curr_depth of this vertex in evaluate()
        depth = depth + 1;
        if (depth > cx.counters->max_depth) cx.counters->max_depth = depth;
        if (!expand(depth)) break;
        bool alive = false;
        Ray next_ray{};
        Color Tn = T;
        float bsdf_pdf = 0.0f;
        bool mis_next = true;
        {
            SampledDirection sb;
            P2 s2 = sampler.next2d();
            if (bsdf.sample(cx.math, its.uv, its.wi, s2, &sb)) {
                V3 d_out_global = its.frame.to_world(sb.d);
                Tn = T * sb.weight;
                if (!Tn.is_zero()) {
                    bool do_rr = I.rr_depth < 0 ? true : (uint32_t)I.rr_depth <= depth;
                    bool survive = true;
                    float rr_weight = 1.0f;
                    if (do_rr) {
                        float q = rmin(Tn.channel_max(), 0.95f);
                        if (q < sampler.next()) survive = false;
                        else rr_weight = 1.0f / q;
                    }
                    if (survive) {
                        Tn.scale(rr_weight);
                        next_ray = spawn_ray(its, d_out_global);
                        bsdf_pdf = sb.pdf.value();
                        mis_next = sb.pdf.kind == PDF::SolidAngle && !bsdf.is_smooth();
                        alive = true;
                    }
                }
            }
        }
        // ---- light sampling strategy (runs even if the bounce died) ----
        if (use_nee && !bsdf.is_smooth()) {
            float r_sel = sampler.next();
            float r = sampler.next();
            P2 uv = sampler.next2d();
            LightSampling rec = sc.emitters.sample_light(cx.math, its.p, &its.n_s, r_sel, r, uv);
            bool visible = sc.visible(its.p, rec.p, cx.accel_mode, *cx.counters);
            if (rec.is_valid() && !mute && add_ok(vdepth) && I.strategy != RL_STRATEGY_BSDF) {
                V3 wo = its.frame.to_local(rec.d);
                Color f = bsdf.eval(cx.math, its.uv, its.wi, wo);
                Color contrib = T * (rec.weight * f);
                if (!contrib.is_zero()) {
                    float w = 1.0f;
                    if (I.strategy == RL_STRATEGY_ALL && rec.pdf.kind == PDF::SolidAngle) { // a Discrete light edge has no MIS (path.rs:80)
                        // the graph evaluates the BSDF pdf along Edge::from_vertex's direction (p_light - p) / |.| (edge.rs:37-39):
                        // bit-identical to rec.d for mesh lights, not for the environment (p = x + d t)
                        V3 de = rec.p - its.p;
                        de = de / magnitude(de);
                        float pb = bsdf.pdf(cx.math, its.uv, its.wi, rec.emitter == sc.env_emitter ? its.frame.to_local(de) : wo).value();
                        float pl = rec.pdf.value();
                        w = pl / (pb + pl);
                    }
                    if (visible) {
                        L = L + contrib * w;
                        cx.counters->nee_added++;
                    }
                }
            }
        }
        if (!alive) break;
        T = Tn;
        pdf_prev = bsdf_pdf;
        mis_prev = mis_next;
        ray = next_ray;
    }
    return L;
}

// IntegratorDirect::compute_pixel, direct.rs:21-233 (no environment map)
Color direct_compute_pixel(const rl_integrator_desc &I, uint32_t ix, uint32_t iy, const Ctx &cx, Sampler &sampler) {
    const Scene &sc = *cx.scene;
    float jx = sampler.next();
    float jy = sampler.next();
    Ray ray = sc.camera.generate(P2{(float)ix + jx, (float)iy + jy});
    Color l_i = Color::zero();
    if (1 > cx.counters->max_depth) cx.counters->max_depth = 1;
    Intersection its;
    if (!sc.trace(ray, cx.accel_mode, *cx.counters, &its)) return sc.enviroment_luminance(cx.math, ray.d); // direct.rs:33-36
    if (its.cos_theta() <= 0.0f) return l_i;
    l_i = l_i + its.mesh->emit(its.uv);
    float weight_nb_bsdf = I.nb_bsdf_samples == 0 ? 0.0f : 1.0f / (float)I.nb_bsdf_samples;
    float weight_nb_light = I.nb_light_samples == 0 ? 0.0f : 1.0f / (float)I.nb_light_samples;
    const BSDF &bsdf = *its.mesh->bsdf;
    for (uint32_t i = 0; i < I.nb_light_samples; i++) { // :63-129
        float r_sel = sampler.next();
        float r = sampler.next();
        P2 uv = sampler.next2d();
        LightSampling rec = sc.emitters.sample_light(cx.math, its.p, &its.n_s, r_sel, r, uv);
        V3 d_out_local = its.frame.to_local(rec.d);
        if (rec.is_valid() && sc.visible(its.p, rec.p, cx.accel_mode, *cx.counters) && !bsdf.is_smooth()) {
            float pdf_bsdf = bsdf.pdf(cx.math, its.uv, its.wi, d_out_local).value();
            float weight_light = rec.pdf.kind == PDF::Discrete ? 1.0f // (PDF::Discrete(_), _) => 1.0, direct.rs:110
                                                               : mis_weight(rec.pdf.value() * weight_nb_light, pdf_bsdf * weight_nb_bsdf);
            l_i = l_i + weight_light * bsdf.eval(cx.math, its.uv, its.wi, d_out_local) * weight_nb_light * rec.weight;
        }
    }
    for (uint32_t i = 0; i < I.nb_bsdf_samples; i++) { // :135-230
        SampledDirection sb;
        P2 s2 = sampler.next2d();
        if (!bsdf.sample(cx.math, its.uv, its.wi, s2, &sb)) continue;
        V3 d_out_world = its.frame.to_world(sb.d);
        Ray r2 = spawn_ray(its, d_out_world);
        Intersection next_its;
        if (sc.trace(r2, cx.accel_mode, *cx.counters, &next_its)) {
            if (next_its.mesh->is_light() && dot(next_its.n_g, -r2.d) > 0.0f) {
                float weight_bsdf = 1.0f; // PDF::Discrete(_v) => 1.0 (direct.rs:170)
                if (sb.pdf.kind == PDF::SolidAngle) {
                    float light_pdf = sc.emitters.direct_pdf(next_its.mesh, LightSamplingPDF{r2.o, next_its.p, next_its.n_g, r2.d}, &its.n_s, (long)next_its.primitive_id).value();
                    weight_bsdf = mis_weight(sb.pdf.value() * weight_nb_bsdf, light_pdf * weight_nb_light);
                }
                l_i = l_i + weight_bsdf * sb.weight * next_its.mesh->emit(next_its.uv) * weight_nb_bsdf;
            }
        } else if (sc.has_environment) { // direct.rs:183-227
            float weight_bsdf = 1.0f;
            if (sb.pdf.kind == PDF::SolidAngle) {
                float t;
                if (!bsphere_intersect(sc.env_emitter->bsphere, r2, &t)) std::abort(); // t.unwrap()
                V3 p = r2.o + r2.d * t;
                V3 n = normalize(sc.env_emitter->bsphere.center - p);
                float light_pdf = sc.emitters.direct_pdf(sc.env_emitter, LightSamplingPDF{r2.o, p, n, r2.d, cx.math.mode}).value();
                weight_bsdf = mis_weight(sb.pdf.value() * weight_nb_bsdf, light_pdf * weight_nb_light);
            }
            l_i = l_i + weight_bsdf * sb.weight * sc.enviroment_luminance(cx.math, r2.d) * weight_nb_bsdf;
        }
    }
    return l_i;
}

// IntegratorAO::compute_pixel, ao.rs:20-72
Color ao_compute_pixel(const rl_integrator_desc &I, uint32_t ix, uint32_t iy, const Ctx &cx, Sampler &sampler) {
    const Scene &sc = *cx.scene;
    float jx = sampler.next();
    float jy = sampler.next();
    Ray ray = sc.camera.generate(P2{(float)ix + jx, (float)iy + jy});
    if (1 > cx.counters->max_depth) cx.counters->max_depth = 1;
    Intersection its;
    if (!sc.trace(ray, cx.accel_mode, *cx.counters, &its)) return Color::zero();
    const bool normal_correction = I.ao_normal_correction != 0;
    if (!normal_correction && its.cos_theta() <= 0.0f) return Color::zero();
    bool flipped = normal_correction && its.cos_theta() <= 0.0f;
    V3 d_local = cosine_sample_hemisphere(cx.math, sampler.next2d());
    V3 d_world = flipped ? its.frame.to_world(-d_local) : its.frame.to_world(d_local);
    Ray r2 = spawn_ray(its, d_world);
    Intersection new_its;
    if (!sc.trace(r2, cx.accel_mode, *cx.counters, &new_its)) return Color::one();
    if (I.ao_max_distance < 0.0f) return Color::zero(); // max_distance: None
    return new_its.dist > I.ao_max_distance ? Color::one() : Color::zero();
}

Color compute_pixel(const rl_integrator_desc &I, uint32_t estimator, uint32_t ix, uint32_t iy, const Ctx &cx, Sampler &sampler) {
    if (I.kind == RL_INTEGRATOR_AO) return ao_compute_pixel(I, ix, iy, cx, sampler);
    if (I.kind == RL_INTEGRATOR_DIRECT) return direct_compute_pixel(I, ix, iy, cx, sampler);
    if (estimator == ORC_EST_STREAM) return path_compute_pixel_stream(I, ix, iy, cx, sampler);
    return path_compute_pixel(I, ix, iy, cx, sampler);
}

// Image-tile ownership shared with the GPU library: 16x16 tiles,
// tile (tx,ty) belongs to rank (tx + ty) % nranks.
inline bool tile_owned(uint32_t tx, uint32_t ty, uint32_t rank, uint32_t nranks) { return nranks <= 1 || ((tx + ty) % nranks) == rank; }

} // namespace

// ================================================================================================
// C entry points
// ================================================================================================
struct orc_scene {
    Scene scene;
};

extern "C" {

orc_scene *orc_scene_create(const rl_scene_desc *desc, char *err, size_t errlen) {
    auto fail = [&](const char *m) -> orc_scene * {
        if (err && errlen) {
            std::strncpy(err, m, errlen - 1);
            err[errlen - 1] = 0;
        }
        return nullptr;
    };
    if (!desc || !desc->meshes || desc->nmeshes == 0) return fail("empty scene");
    if (desc->has_volume || desc->has_environment > 2) return fail("volumes are outside the hot path");
    if (desc->has_environment == 2 && (desc->environment_texture == 0 || desc->environment_texture > desc->ntextures || !desc->textures ||
                                        desc->textures[desc->environment_texture - 1].kind != RL_TEX_BITMAP))
        return fail("has_environment == 2 needs environment_texture = 1 + index of a bitmap texture");
    auto *os = new orc_scene;
    Scene &s = os->scene;
    s.camera.img_x = desc->camera.width, s.camera.img_y = desc->camera.height;
    std::memcpy(s.camera.sample_to_camera.m, desc->camera.sample_to_camera, 64);
    std::memcpy(s.camera.to_world.m, desc->camera.to_world, 64);
    uint32_t first = 0;
    for (uint32_t i = 0; i < desc->nmeshes; i++) {
        const rl_mesh_desc &md = desc->meshes[i];
        auto m = std::make_unique<Mesh>();
        for (uint32_t v = 0; v < md.nverts; v++) m->vertices.push_back(load3(md.P + 3 * v));
        for (uint32_t t = 0; t < md.ntris; t++) m->indices.push_back(Idx3{md.idx[3 * t], md.idx[3 * t + 1], md.idx[3 * t + 2]});
        if (md.N) {
            m->has_normals = true;
            for (uint32_t v = 0; v < md.nverts; v++) m->normals.push_back(load3(md.N + 3 * v));
        }
        if (md.UV) {
            m->has_uv = true;
            for (uint32_t v = 0; v < md.nverts; v++) m->uv.push_back(P2{md.UV[2 * v], md.UV[2 * v + 1]});
        }
        if (md.mat.kd_texture > desc->ntextures || md.mat.ks_texture > desc->ntextures || md.mat.kt_texture > desc->ntextures ||
            md.mat.eta_texture > desc->ntextures || md.mat.k_texture > desc->ntextures)
            return fail("texture index out of range");
        try {
            m->bsdf = make_bsdf(md.mat, desc->textures, desc->ntextures, desc->submaterials, desc->nsubmaterials);
        } catch (const std::exception &e) {
            delete os;
            return fail(e.what());
        }
        m->light = md.emission_kind != 0;
        m->emission = Color{md.emission[0], md.emission[1], md.emission[2]};
        if (md.emission_kind > RL_EMISSION_TEXTURE) return fail("unknown emission kind");
        if (md.emission_kind >= RL_EMISSION_HSV) { // EmissionType::HSV { scale } | Texture { scale, img }
            if (!md.UV) return fail("HSV / textured emission needs uv coordinates (uv.unwrap(), geometry.rs:197)");
            m->emission_type = md.emission_kind == RL_EMISSION_HSV ? Mesh::EHsv : Mesh::ETexture;
            m->emission_scale = md.emission[0];
            if (md.emission_kind == RL_EMISSION_TEXTURE) {
                if (md.emission_texture == 0 || md.emission_texture > desc->ntextures || desc->textures[md.emission_texture - 1].kind != RL_TEX_BITMAP)
                    return fail("emission_texture must be 1 + index of a bitmap texture");
                const rl_texture &t = desc->textures[md.emission_texture - 1];
                m->emission_img = std::make_shared<BitmapTex>();
                m->emission_img->size_x = t.width, m->emission_img->size_y = t.height;
                for (size_t i = 0; i < (size_t)t.width * t.height; i++) m->emission_img->colors.push_back(Color{t.pixels[3 * i], t.pixels[3 * i + 1], t.pixels[3 * i + 2]});
            }
        }
        m->first_prim = first;
        first += md.ntris;
        m->build_cdf();
        s.meshes.push_back(std::move(m));
    }
    if (desc->has_environment) s.has_environment = true, s.environment = Color{desc->environment[0], desc->environment[1], desc->environment[2]};
    if (desc->has_environment == 2) { // EnvironmentLightColor::new_texture(image)
        const rl_texture &t = desc->textures[desc->environment_texture - 1];
        s.environment_image.size_x = t.width, s.environment_image.size_y = t.height;
        for (size_t i = 0; i < (size_t)t.width * t.height; i++) s.environment_image.colors.push_back(Color{t.pixels[3 * i], t.pixels[3 * i + 1], t.pixels[3 * i + 2]});
    }
    for (uint32_t li = 0; li < desc->nlights; li++) {
        if (desc->lights[li].kind > RL_LIGHT_DIRECTIONAL) return fail("unknown light kind");
        s.lights.push_back(desc->lights[li]);
    }
    s.build_ats = desc->use_ats != 0;
    s.build_emitters();
    if (!s.ats_error.empty()) {
        std::string m = s.ats_error;
        delete os;
        return fail(m.c_str());
    }
    s.build_bvh();
    return os;
}
void orc_scene_destroy(orc_scene *s) { delete s; }

void orc_bvh_info(const orc_scene *s, uint32_t *nnodes, uint32_t *nprims, float root_min[3], float root_max[3]) {
    if (nnodes) *nnodes = (uint32_t)s->scene.nodes.size();
    if (nprims) *nprims = (uint32_t)s->scene.primitives.size();
    if (root_min) store3(root_min, s->scene.nodes[0].aabb.p_min);
    if (root_max) store3(root_max, s->scene.nodes[0].aabb.p_max);
}

int orc_render(const orc_scene *os, const rl_integrator_desc *integ, uint32_t spp, uint64_t seed, uint32_t sampler_mode,
               const orc_config *cfg, float *out_rgb, orc_stats *stats) {
    if (!os || !integ || !cfg || !out_rgb || spp == 0) return RL_ERR_INVALID; // assert_ne!(nb_samples, 0), mod.rs:410
    const Scene &sc = os->scene;
    const uint32_t W = sc.camera.img_x, H = sc.camera.img_y;
    if (integ->kind == RL_INTEGRATOR_PATH) {
        // max_depth < 2: the sensor vertex is never expanded and evaluate() unwraps a None edge (path.rs:154) -> panic
        if (integ->max_depth >= 0 && integ->max_depth < 2) return RL_ERR_INVALID;
        if (sc.emitters.emitters.empty() && integ->strategy != RL_STRATEGY_BSDF) return RL_ERR_INVALID; // scene.rs:97-100
    } else if (integ->kind == RL_INTEGRATOR_DIRECT) {
        if (sc.emitters.emitters.empty() && integ->nb_light_samples > 0) return RL_ERR_INVALID;
    } else if (integ->kind == RL_INTEGRATOR_AO) {
        if (integ->ao_max_distance != integ->ao_max_distance) return RL_ERR_INVALID;
    } else return RL_ERR_INVALID;
    // generate_img_blocks, mod.rs:351-374: x-major block order, one cloned sampler per block
    struct Block {
        uint32_t px, py, sx, sy;
        IndependentSampler sampler;
        std::vector<Color> colors;
    };
    std::vector<Block> blocks;
    IndependentSampler master;
    master.seeding = cfg->seeding;
    master.rnd = Xoshiro256PP::seed_from_u64(seed, cfg->seeding); // cli.rs:886-890
    for (uint32_t ix = 0; ix < W; ix += 16)
        for (uint32_t iy = 0; iy < H; iy += 16) {
            Block b;
            b.px = ix, b.py = iy, b.sx = std::min(16u, W - ix), b.sy = std::min(16u, H - iy);
            b.sampler = master.clone_box();
            blocks.push_back(std::move(b));
        }
    uint32_t nthreads = cfg->nthreads ? cfg->nthreads : std::max(1u, std::thread::hardware_concurrency());
    nthreads = std::min<uint32_t>(nthreads, (uint32_t)blocks.size());
    std::vector<Counters> counters(nthreads);
    std::atomic<size_t> next_block{0};
    auto t0 = std::chrono::steady_clock::now();
    auto worker = [&](uint32_t tid) {
        Ctx cx{&sc, cfg->accel_mode, Math{cfg->math_mode}, &counters[tid]};
        for (;;) {
            size_t bi = next_block.fetch_add(1);
            if (bi >= blocks.size()) break;
            Block &b = blocks[bi];
            b.colors.assign((size_t)b.sx * b.sy, Color::zero());
            if (!tile_owned(b.px / 16, b.py / 16, cfg->rank, cfg->nranks)) continue;
            for (uint32_t iy = 0; iy < b.sy; iy++)
                for (uint32_t ix = 0; ix < b.sx; ix++) {
                    Color &acc = b.colors[(size_t)iy * b.sx + ix];
                    for (uint32_t s = 0; s < spp; s++) {
                        uint32_t gx = ix + b.px, gy = iy + b.py;
                        Color c;
                        if (sampler_mode == RL_SAMPLER_COUNTER) {
                            CounterSampler cs(seed, gy * W + gx, cfg->sample_offset + s);
                            c = compute_pixel(*integ, cfg->estimator, gx, gy, cx, cs);
                        } else c = compute_pixel(*integ, cfg->estimator, gx, gy, cx, b.sampler);
                        acc = acc + c; // Bitmap::accumulate, structure.rs:397-402
                    }
                }
            float f = 1.0f / (float)spp; // im_block.scale(1/spp), mod.rs:436
            for (Color &c : b.colors) c.scale(f);
        }
    };
    if (nthreads == 1) worker(0);
    else {
        std::vector<std::thread> th;
        for (uint32_t t = 0; t < nthreads; t++) th.emplace_back(worker, t);
        for (auto &t : th) t.join();
    }
    // image.accumulate_bitmap(block), mod.rs:445-449
    std::memset(out_rgb, 0, sizeof(float) * 3 * (size_t)W * H);
    for (auto &b : blocks)
        for (uint32_t y = 0; y < b.sy; y++)
            for (uint32_t x = 0; x < b.sx; x++) {
                size_t idx = (size_t)(b.py + y) * W + (b.px + x);
                const Color &c = b.colors[(size_t)y * b.sx + x];
                out_rgb[3 * idx] += c.r, out_rgb[3 * idx + 1] += c.g, out_rgb[3 * idx + 2] += c.b;
            }
    auto t1 = std::chrono::steady_clock::now();
    if (stats) {
        *stats = orc_stats{};
        for (auto &c : counters) {
            stats->segments += c.segments, stats->shadow_rays += c.shadow_rays, stats->shadow_visible += c.shadow_visible;
            stats->hits += c.hits;
            stats->nee_added += c.nee_added;
            stats->max_depth_seen = std::max<uint64_t>(stats->max_depth_seen, c.max_depth);
        }
        uint64_t npix = 0;
        for (auto &b : blocks)
            if (tile_owned(b.px / 16, b.py / 16, cfg->rank, cfg->nranks)) npix += (uint64_t)b.sx * b.sy;
        stats->samples = npix * spp;
        stats->seconds = std::chrono::duration<double>(t1 - t0).count();
        stats->threads_used = nthreads;
    }
    return RL_OK;
}

void orc_path_sample(const orc_scene *os, const rl_integrator_desc *integ, uint64_t seed, uint32_t px, uint32_t py, uint32_t sample,
                     const orc_config *cfg, float rgb[3], uint32_t *n_segments, uint32_t *n_shadow, uint32_t *n_draws) {
    Counters c;
    Ctx cx{&os->scene, cfg->accel_mode, Math{cfg->math_mode}, &c};
    CounterSampler cs(seed, py * os->scene.camera.img_x + px, sample);
    Color r = compute_pixel(*integ, cfg->estimator, px, py, cx, cs);
    rgb[0] = r.r, rgb[1] = r.g, rgb[2] = r.b;
    if (n_segments) *n_segments = (uint32_t)c.segments;
    if (n_shadow) *n_shadow = (uint32_t)c.shadow_rays;
    if (n_draws) *n_draws = cs.draws;
}

int orc_trace(const orc_scene *os, uint32_t accel_mode, size_t n, const float *o, const float *d, uint32_t *prim, float *tuv,
              float *p, float *n_g, float *n_s, float *wi) {
    const Scene &sc = os->scene;
    for (size_t i = 0; i < n; i++) {
        Ray ray = ray_new(load3(o + 3 * i), load3(d + 3 * i));
        IntersectionUV its;
        TriRef res{0, 0};
        if (!sc.trace_uv(ray, accel_mode, its, &res)) {
            prim[i] = 0xFFFFFFFFu;
            if (tuv) tuv[3 * i] = tuv[3 * i + 1] = tuv[3 * i + 2] = 0.0f;
            continue;
        }
        prim[i] = sc.meshes[res.id_mesh]->first_prim + res.id_tri;
        if (tuv) tuv[3 * i] = its.t, tuv[3 * i + 1] = its.u, tuv[3 * i + 2] = its.v;
        if (p || n_g || n_s || wi) {
            Intersection f = fill_intersection(sc.meshes[res.id_mesh].get(), res.id_tri, its.u, its.v, ray, its.n, its.t, its.p);
            if (p) store3(p + 3 * i, f.p);
            if (n_g) store3(n_g + 3 * i, f.n_g);
            if (n_s) store3(n_s + 3 * i, f.n_s);
            if (wi) store3(wi + 3 * i, f.wi);
        }
    }
    return RL_OK;
}
int orc_visible(const orc_scene *os, uint32_t accel_mode, size_t n, const float *p0, const float *p1, uint8_t *out) {
    Counters c;
    for (size_t i = 0; i < n; i++) out[i] = os->scene.visible(load3(p0 + 3 * i), load3(p1 + 3 * i), accel_mode, c) ? 1 : 0;
    return RL_OK;
}
int orc_primary_hits(const orc_scene *os, uint32_t accel_mode, uint32_t *prim, float *tuv) {
    const Scene &sc = os->scene;
    uint32_t W = sc.camera.img_x, H = sc.camera.img_y;
    for (uint32_t y = 0; y < H; y++)
        for (uint32_t x = 0; x < W; x++) {
            Ray ray = sc.camera.generate(P2{(float)x + 0.5f, (float)y + 0.5f});
            size_t i = (size_t)y * W + x;
            orc_trace(os, accel_mode, 1, &ray.o.x, &ray.d.x, prim + i, tuv ? tuv + 3 * i : nullptr, nullptr, nullptr, nullptr, nullptr);
        }
    return RL_OK;
}

int orc_intersect_tri(const float v0[3], const float v1[3], const float v2[3], const float o[3], const float d[3], float *t_io,
                      float *u, float *v, float p[3], float n[3]) {
    Mesh m;
    m.vertices = {load3(v0), load3(v1), load3(v2)};
    m.indices = {Idx3{0, 1, 2}};
    IntersectionUV its{*t_io, V3{0, 0, 0}, V3{0, 0, 0}, 0, 0};
    bool hit = m.intersection_tri(0, load3(o), load3(d), its);
    if (hit) {
        *t_io = its.t;
        if (u) *u = its.u;
        if (v) *v = its.v;
        if (p) store3(p, its.p);
        if (n) store3(n, its.n);
    }
    return hit ? 1 : 0;
}
int orc_aabb_intersect(const float pmin[3], const float pmax[3], const float o[3], const float d[3], float tnear, float tfar, float *t) {
    AABB a;
    a.p_min = load3(pmin), a.p_max = load3(pmax);
    Ray r{load3(o), load3(d), tnear, tfar};
    float tt;
    if (!a.intersect(r, &tt)) return 0;
    if (t) *t = tt;
    return 1;
}
void orc_frame(const float n[3], float out9[9]) {
    Frame f(load3(n));
    store3(out9, f.x), store3(out9 + 3, f.y), store3(out9 + 6, f.z);
}
void orc_cosine_sample_hemisphere(uint32_t math_mode, float u0, float u1, float out[3]) { store3(out, cosine_sample_hemisphere(Math{math_mode}, P2{u0, u1})); }
void orc_uniform_sample_triangle(float u0, float u1, float out[2]) {
    P2 b = uniform_sample_triangle(P2{u0, u1});
    out[0] = b.x, out[1] = b.y;
}
float orc_dist1d_normalize(const float *elements, uint32_t n, float *cdf) {
    Distribution1D d = Distribution1D::normalize(std::vector<float>(elements, elements + n));
    std::memcpy(cdf, d.cdf.data(), sizeof(float) * (n + 1));
    return d.func_int;
}
uint32_t orc_dist1d_sample_discrete(const float *cdf, uint32_t n_plus_1, float v) {
    Distribution1D d;
    d.cdf.assign(cdf, cdf + n_plus_1);
    return (uint32_t)d.sample_discrete(v);
}
float orc_mis_weight(float a, float b) { return mis_weight(a, b); }
// one material; a blend is passed as an array of three materials {blend, bsdf1, bsdf2} with blend_a = 1, blend_b = 2
static std::unique_ptr<BSDF> make_bsdf1(const rl_material *m) { return make_bsdf(*m, nullptr, 0, m + 1, m->kind == RL_BSDF_BLEND ? 2u : 0u); }
int orc_bsdf_sample(uint32_t math_mode, const rl_material *m, const float wi[3], float s0, float s1, float weight[3], float d[3], float *pdf) {
    auto b = make_bsdf1(m);
    SampledDirection sd;
    if (!b->sample(Math{math_mode}, UV{}, load3(wi), P2{s0, s1}, &sd)) return 0;
    weight[0] = sd.weight.r, weight[1] = sd.weight.g, weight[2] = sd.weight.b;
    store3(d, sd.d);
    *pdf = sd.pdf.value();
    return sd.pdf.kind == PDF::Discrete ? 2 : 1; // 2: PDF::Discrete
}
int orc_bsdf_flags(const rl_material *m) { // bit 0: is_twosided, bit 1: bsdf_type().is_smooth()
    auto b = make_bsdf1(m);
    return (b->is_twosided() ? 1 : 0) | (b->is_smooth() ? 2 : 0);
}
float orc_bsdf_pdf(uint32_t math_mode, const rl_material *m, const float wi[3], const float wo[3]) {
    return make_bsdf1(m)->pdf(Math{math_mode}, UV{}, load3(wi), load3(wo)).value();
}
void orc_bsdf_eval(uint32_t math_mode, const rl_material *m, const float wi[3], const float wo[3], float out[3]) {
    Color c = make_bsdf1(m)->eval(Math{math_mode}, UV{}, load3(wi), load3(wo));
    out[0] = c.r, out[1] = c.g, out[2] = c.b;
}
int orc_sample_light(const orc_scene *os, const float x[3], float r_sel, float r, float u0, float u1, float p[3], float n[3], float d[3],
                     float weight[3], float *pdf) {
    const Scene &sc = os->scene;
    if (sc.emitters.emitters.empty()) return -1;
    LightSampling rec = sc.emitters.sample_light(Math{ORC_MATH_SPEC}, load3(x), nullptr, r_sel, r, P2{u0, u1});
    store3(p, rec.p), store3(n, rec.n), store3(d, rec.d);
    weight[0] = rec.weight.r, weight[1] = rec.weight.g, weight[2] = rec.weight.b;
    *pdf = rec.pdf.value();
    for (size_t i = 0; i < sc.meshes.size(); i++)
        if (sc.emitters.of_mesh(sc.meshes[i].get()) == rec.emitter) return (int)i;
    return -2 - (rec.pdf.kind == PDF::Discrete ? 1 : 0); // a non-mesh emitter (-3: PDF::Discrete)
}
float orc_direct_pdf(const orc_scene *os, uint32_t mesh, const float o[3], const float p[3], const float n[3], const float dir[3]) {
    const Scene &sc = os->scene;
    return sc.emitters.direct_pdf(sc.meshes[mesh].get(), LightSamplingPDF{load3(o), load3(p), load3(n), load3(dir)}).value();
}
int orc_camera_new(uint32_t w, uint32_t h, int fov_axis, float fov_deg, const float to_world[16], int flip, float sample_to_camera[16]) {
    M4 tw, s2c;
    std::memcpy(tw.m, to_world, 64);
    if (!camera_new(w, h, fov_axis, fov_deg, tw, flip != 0, &s2c)) return -1;
    std::memcpy(sample_to_camera, s2c.m, 64);
    return 0;
}
void orc_camera_generate(const orc_scene *os, float px, float py, float o[3], float d[3]) {
    Ray r = os->scene.camera.generate(P2{px, py});
    store3(o, r.o), store3(d, r.d);
}
void orc_sampler_block_stream(uint64_t seed, uint32_t seeding, uint32_t block, uint32_t n, float *out) {
    IndependentSampler master;
    master.seeding = seeding;
    master.rnd = Xoshiro256PP::seed_from_u64(seed, seeding);
    IndependentSampler b = master.clone_box();
    for (uint32_t i = 0; i < block; i++) b = master.clone_box();
    for (uint32_t i = 0; i < n; i++) out[i] = b.next();
}
void orc_sampler_counter(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t n, float *out) {
    CounterSampler cs(seed, pixel, sample);
    for (uint32_t i = 0; i < n; i++) out[i] = cs.next();
}
uint64_t orc_xoshiro_next_u64(uint64_t state[4]) {
    Xoshiro256PP x;
    std::memcpy(x.s, state, 32);
    uint64_t r = x.next_u64();
    std::memcpy(state, x.s, 32);
    return r;
}
void orc_spec_sincos(float x, float *s, float *c) { spec_sincos(x, s, c); }
static bool ats_locate(const Scene &sc, uint32_t prim, size_t *emitter, size_t *tri) { // global triangle index -> (emitter id, triangle of its mesh)
    for (size_t e = 0; e < sc.emitters.emitters.size(); e++) {
        const Mesh *m = sc.emitters.emitter_mesh[e];
        if (m && prim >= m->first_prim && prim < m->first_prim + m->indices.size()) {
            *emitter = e, *tri = prim - m->first_prim;
            return true;
        }
    }
    return false;
}
int orc_ats_sample(const orc_scene *os, float r, const float p[3], const float n[3], int has_n, uint32_t *prim, float *pdf) {
    const Scene &sc = os->scene;
    if (!sc.emitters.ats) return -1;
    V3 nn = load3(n);
    const LightProxy &lp = sc.emitters.ats->sample(r, load3(p), has_n ? &nn : nullptr, pdf);
    *prim = (uint32_t)(sc.emitters.emitter_mesh[lp.emitter_id]->first_prim + lp.primitive_idx);
    return 0;
}
int orc_ats_pdf(const orc_scene *os, uint32_t prim, const float p[3], const float n[3], int has_n, float *pdf) {
    const Scene &sc = os->scene;
    size_t e, t;
    if (!sc.emitters.ats || !ats_locate(sc, prim, &e, &t)) return -1;
    V3 nn = load3(n);
    *pdf = sc.emitters.ats->pdf(e, t, load3(p), has_n ? &nn : nullptr);
    return 0;
}
float orc_spec_atan2(float y, float x) { return spec_atan2(y, x); }
float orc_spec_acos(float x) { return spec_acos(x); }
int orc_env_eval_pdf(const orc_scene *os, uint32_t math_mode, const float d[3], float rgb[3], float *pdf) {
    const EnvironmentLight *env = os->scene.env_emitter;
    if (!env) return -1;
    Color c = env->luminance.eval(Math{math_mode}, load3(d));
    rgb[0] = c.r, rgb[1] = c.g, rgb[2] = c.b;
    *pdf = env->luminance.pdf(Math{math_mode}, load3(d));
    return 0;
}
int orc_env_sample(const orc_scene *os, uint32_t math_mode, float u0, float u1, float d[3], float rgb[3], float *pdf) {
    const EnvironmentLight *env = os->scene.env_emitter;
    if (!env) return -1;
    V3 dd;
    Color c;
    env->luminance.sample_direction(Math{math_mode}, P2{u0, u1}, &dd, &c, pdf);
    store3(d, dd);
    rgb[0] = c.r, rgb[1] = c.g, rgb[2] = c.b;
    return 0;
}
float orc_spec_powf(float x, float y) { return spec_powf(x, y); }

} // extern "C"
