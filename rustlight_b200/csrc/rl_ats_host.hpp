// rl_ats_host.hpp -- the light tree of `-x ats` (LightSamplerATS, src/emitter.rs:782-1488), built on the host once per scene.
//
// Scene::build_emitters(true) -> EmitterSampler::build_ats (emitter.rs:1505-1508): every emitter must be a surface (:1292-1294), each
// triangle of an emissive mesh becomes a LightProxy (Mesh::convert_light_proxy, :730-779: orientation cone theta_o = 0, theta_e = pi/2
// around the un-flipped geometric normal, phi = max channel of Le x area), and build_bvh (:1145-1287) splits the proxies recursively
// over 12 centroid buckets per axis with the cost  kr (phi_0 M(b0) A(b0) + phi_1 M(b1) A(b1))  -- splits with an empty side cost
// NaN (0 x inf) and are never taken; no valid split -> the middle of the slice.  Node bounds are merged with LightBounds::union /
// DirectionCone::union (:857-973), whose acos / asin / sin / cos run here with the host's libm, as they do in the reference (the oracle
// carries an independently typed copy and runs on the same host).
//
// Device layout (rl_device.cuh: ats_*): four float4 per node
//   [0] {aabb centre, bounding-sphere radius}   [1] {w, phi}   [2] {cos_theta_o, cos_theta_e, two_sided, -}
//   [3] {left, right, parent, light} as uint bits (0xffffffff = none; light = global triangle index of the leaf's proxy)
// plus leaf_of_prim[global triangle] (0xffffffff for triangles that emit nothing) for the pdf walk from a leaf to the root.
// itertools::partition (the in-place partition of :1257-1263) is restated from its published algorithm: front cursor advances over
// elements that satisfy the predicate, a failing one is swapped with the last satisfying element found from the back.
#pragma once
#include <cmath>
#include <cstdint>
#include <string>
#include <vector>

#include "rl_device.cuh"

namespace rl {

struct AtsBounds { // LightBounds, emitter.rs:902-937
    float lo[3], hi[3]; // aabb (default: f32::MAX / f32::MIN)
    float w[3] = {0.0f, 0.0f, 1.0f};
    float phi = 0.0f, theta_o = 0.0f, theta_e = 0.0f, cos_theta_o = 1.0f, cos_theta_e = 1.0f;
    bool two_sided = false;
    AtsBounds() {
        for (int a = 0; a < 3; a++) lo[a] = RL_F32_MAX, hi[a] = -RL_F32_MAX;
    }
};
struct AtsProxy {
    uint32_t prim; // global triangle index (mesh-major)
    AtsBounds b;
};
struct AtsTree {
    std::vector<float4> nodes;         // 4 per node
    std::vector<uint32_t> leaf_of_prim; // ntris entries
    uint32_t root = 0, depth = 0;
};

namespace ats_detail {
inline float safe_acos(float v) { return std::acos(std::fmin(std::fmax(v, -1.0f), 1.0f)); }
inline float safe_asin(float v) { return std::asin(std::fmin(std::fmax(v, -1.0f), 1.0f)); }
inline float len3(const float *v) { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
inline float angle_between(const float *a, const float *b) { // :796-802
    const float d = a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
    if (d < 0.0f) {
        const float s[3] = {b[0] + a[0], b[1] + a[1], b[2] + a[2]};
        return 3.14159265358979323846f - 2.0f * safe_asin(len3(s) / 2.0f);
    }
    const float s[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]};
    return 2.0f * safe_asin(len3(s) / 2.0f);
}
struct Cone {
    float w[3], cos_theta;
};
inline Cone entire_sphere() { return Cone{{0.0f, 0.0f, 1.0f}, -1.0f}; }
inline Cone cone_union(const Cone &a, const Cone &b) { // DirectionCone::union, :857-899 (neither cone is ever `empty` here)
    const float PI_F = 3.14159265358979323846f;
    const float theta_a = safe_acos(a.cos_theta), theta_b = safe_acos(b.cos_theta), theta_d = angle_between(a.w, b.w);
    if (std::fmin(theta_d + theta_b, PI_F) <= theta_a) return a;
    if (std::fmin(theta_d + theta_a, PI_F) <= theta_b) return b;
    const float theta_o = (theta_a + theta_d + theta_b) / 2.0f;
    if (theta_o >= PI_F) return entire_sphere();
    const float theta_r = theta_o - theta_a;
    const float wr[3] = {a.w[1] * b.w[2] - a.w[2] * b.w[1], a.w[2] * b.w[0] - a.w[0] * b.w[2], a.w[0] * b.w[1] - a.w[1] * b.w[0]};
    if (wr[0] * wr[0] + wr[1] * wr[1] + wr[2] * wr[2] == 0.0f) return entire_sphere();
    // rotate_angle_axis(theta_r.to_degrees(), &wr) (:804-829): degrees and back, then the axis-angle matrix applied to a.w
    // f32::to_degrees multiplies by the literal 57.29577951308232..., f32::to_radians by (PI / 180.0) evaluated in f32
    const float deg = theta_r * 57.2957795130823208767981548141051703f, rad = deg * (PI_F / 180.0f);
    const float st = std::sin(rad), ct = std::cos(rad);
    const float il = 1.0f / len3(wr);
    const float x = wr[0] * il, y = wr[1] * il, z = wr[2] * il;
    // the reference fills Matrix4::new(c0r0, c0r1, c0r2, 0, c1r0, ...) column by column and transposes: row i of the product uses c{i}r{0..2}
    const float r0[3] = {x * x + (1.0f - x * x) * ct, x * y * (1.0f - ct) - z * st, x * z * (1.0f - ct) + y * st};
    const float r1[3] = {x * y * (1.0f - ct) + z * st, y * y + (1.0f - y * y) * ct, y * z * (1.0f - ct) - x * st};
    const float r2[3] = {x * z * (1.0f - ct) - y * st, y * z * (1.0f - ct) + x * st, z * z + (1.0f - z * z) * ct};
    Cone c;
    // Matrix4 * Vector4 in cgmath: c0 * v.x + c1 * v.y + c2 * v.z (+ c3 * 0); after the transpose column j holds (r0[j], r1[j], r2[j])
    c.w[0] = r0[0] * a.w[0] + r0[1] * a.w[1] + r0[2] * a.w[2] + 0.0f; // (+ c3 * 0)
    c.w[1] = r1[0] * a.w[0] + r1[1] * a.w[1] + r1[2] * a.w[2] + 0.0f;
    c.w[2] = r2[0] * a.w[0] + r2[1] * a.w[1] + r2[2] * a.w[2] + 0.0f;
    c.cos_theta = std::cos(theta_o);
    return c;
}
inline AtsBounds bounds_union(const AtsBounds &a, const AtsBounds &b) { // LightBounds::union, :948-973
    if (a.phi == 0.0f) return b;
    if (b.phi == 0.0f) return a;
    const Cone c = cone_union(Cone{{a.w[0], a.w[1], a.w[2]}, a.cos_theta_o}, Cone{{b.w[0], b.w[1], b.w[2]}, b.cos_theta_o});
    AtsBounds r;
    for (int k = 0; k < 3; k++) r.lo[k] = std::fmin(a.lo[k], b.lo[k]), r.hi[k] = std::fmax(a.hi[k], b.hi[k]), r.w[k] = c.w[k];
    r.phi = a.phi + b.phi;
    r.theta_o = safe_acos(c.cos_theta);
    r.theta_e = std::fmax(a.theta_e, b.theta_e);
    r.cos_theta_o = std::cos(r.theta_o);
    r.cos_theta_e = std::cos(r.theta_e);
    r.two_sided = a.two_sided || b.two_sided;
    return r;
}
inline float center_of(const AtsBounds &b, int k) { return (b.hi[k] - b.lo[k]) * 0.5f + b.lo[k]; } // AABB::center
inline float half_area(const float *lo, const float *hi) { // AABB::surface_area (half the true area), structure.rs:822-836
    const float d[3] = {hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]};
    float s = 0.0f;
    for (int i = 0; i < 3; i++) {
        float v = 1.0f;
        for (int j = 0; j < 3; j++)
            if (i != j) v *= d[j];
        s += v;
    }
    return s;
}
inline float momega(const AtsBounds &b) { // :1208-1216
    const float PI_F = 3.14159265358979323846f, FRAC_PI_2 = 1.57079632679489661923f;
    const float theta_w = std::fmin(b.theta_o + b.theta_e, PI_F);
    return 2.0f * PI_F * (1.0f - std::cos(b.theta_o)) +
           FRAC_PI_2 * (2.0f * theta_w * std::sin(b.theta_o) - std::cos(b.theta_o - 2.0f * theta_w) - 2.0f * b.theta_o * std::sin(b.theta_o) + std::cos(b.theta_o));
}
struct Builder {
    std::vector<AtsBounds> node_bounds;
    std::vector<uint32_t> left, right, parent, light, depth_of;
    uint32_t build(AtsProxy *lights, size_t n, uint32_t depth, bool *ok) { // build_bvh, :1145-1287 (`index` is only the proxy's slot: not needed here)
        if (n == 0) { // `[] => unimplemented!()`
            *ok = false;
            return 0;
        }
        if (n == 1) {
            node_bounds.push_back(lights[0].b), left.push_back(0xffffffffu), right.push_back(0xffffffffu), parent.push_back(0xffffffffu), light.push_back(lights[0].prim);
            depth_of.push_back(depth);
            return (uint32_t)node_bounds.size() - 1u;
        }
        float blo[3], bhi[3], clo[3], chi[3];
        for (int k = 0; k < 3; k++) blo[k] = clo[k] = RL_F32_MAX, bhi[k] = chi[k] = -RL_F32_MAX;
        for (size_t i = 0; i < n; i++)
            for (int k = 0; k < 3; k++) {
                blo[k] = std::fmin(blo[k], lights[i].b.lo[k]), bhi[k] = std::fmax(bhi[k], lights[i].b.hi[k]);
                const float c = center_of(lights[i].b, k);
                clo[k] = std::fmin(clo[k], c), chi[k] = std::fmax(chi[k], c);
            }
        const int NB = 12;
        auto bucket = [&](const AtsProxy &l, int dim) { // (NBUCKETS as f32 * centroid_bounds.offset(&pc)[dim]) as usize, clamped
            const float o = center_of(l.b, dim) - clo[dim], s = chi[dim] - clo[dim];
            const float off = s != 0.0f ? o / s : 0.0f;
            const float v = (float)NB * off;
            uint64_t i = !(v > 0.0f) ? 0ull : (v >= 18446744073709551616.0f ? ~0ull : (uint64_t)v);
            return (int)(i < (uint64_t)(NB - 1) ? i : (uint64_t)(NB - 1));
        };
        float min_cost = RL_F32_MAX;
        int min_bucket = -1, min_dim = -1;
        for (int dim = 0; dim < 3; dim++) {
            if (chi[dim] == clo[dim]) continue;
            AtsBounds bb[NB];
            for (size_t i = 0; i < n; i++) {
                const int bi = bucket(lights[i], dim);
                bb[bi] = bounds_union(bb[bi], lights[i].b);
            }
            const float size[3] = {bhi[0] - blo[0], bhi[1] - blo[1], bhi[2] - blo[2]};
            for (int i = 0; i + 1 < NB; i++) {
                AtsBounds b0, b1;
                for (int j = 0; j <= i; j++) b0 = bounds_union(b0, bb[j]);
                for (int j = i + 1; j < NB; j++) b1 = bounds_union(b1, bb[j]);
                const float kr = std::fmax(std::fmax(size[0], size[1]), size[2]) / size[dim];
                const float c = kr * (b0.phi * momega(b0) * half_area(b0.lo, b0.hi) + b1.phi * momega(b1) * half_area(b1.lo, b1.hi));
                if (c > 0.0f && c < min_cost) min_cost = c, min_bucket = i, min_dim = dim;
            }
        }
        size_t mid;
        if (min_dim == -1) mid = n / 2;
        else { // itertools::partition
            size_t split = 0, front = 0, back = n;
            auto pred = [&](const AtsProxy &l) { return bucket(l, min_dim) <= min_bucket; };
            bool done = false;
            while (!done && front < back) {
                if (!pred(lights[front])) {
                    for (;;) {
                        if (back - 1 <= front) {
                            done = true;
                            break;
                        }
                        back--;
                        if (pred(lights[back])) {
                            std::swap(lights[front], lights[back]);
                            break;
                        }
                    }
                    if (done) break;
                }
                front++;
                split++;
            }
            mid = split;
        }
        const uint32_t l = build(lights, mid, depth + 1, ok);
        if (!*ok) return 0;
        const uint32_t r = build(lights + mid, n - mid, depth + 1, ok);
        if (!*ok) return 0;
        node_bounds.push_back(bounds_union(node_bounds[l], node_bounds[r]));
        left.push_back(l), right.push_back(r), parent.push_back(0xffffffffu), light.push_back(0xffffffffu), depth_of.push_back(depth);
        const uint32_t id = (uint32_t)node_bounds.size() - 1u;
        parent[l] = id, parent[r] = id;
        return id;
    }
};
} // namespace ats_detail

// `emissive[p]`: triangle p belongs to an emissive mesh, with emission `le[p]` (three floats).  False + err when the reference would
// panic (no light at all, or a split with an empty side).
inline bool build_ats_tree(const std::vector<float4> &verts, uint32_t ntris, const std::vector<uint8_t> &emissive, const std::vector<float> &le, AtsTree &out, std::string &err) {
    using namespace ats_detail;
    std::vector<AtsProxy> lights;
    for (uint32_t p = 0; p < ntris; p++) {
        if (!emissive[p]) continue;
        const float4 a = verts[3 * (size_t)p], b = verts[3 * (size_t)p + 1], c = verts[3 * (size_t)p + 2];
        const float e1[3] = {b.x - a.x, b.y - a.y, b.z - a.z}, e2[3] = {c.x - a.x, c.y - a.y, c.z - a.z};
        const float n[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]}; // (v1 - v0) x (v2 - v0)
        const float nl = len3(n), il = 1.0f / nl;
        AtsProxy l;
        l.prim = p;
        for (int k = 0; k < 3; k++) l.b.w[k] = n[k] * il; // normalize = v * (1 / |v|)
        l.b.theta_o = 0.0f, l.b.theta_e = 1.57079632679489661923f;
        l.b.phi = std::fmax(le[3 * p], std::fmax(le[3 * p + 1], le[3 * p + 2])) * nl * 0.5f; // emit(uv).channel_max() * n.magnitude() * 0.5
        const float vs[3][3] = {{a.x, a.y, a.z}, {b.x, b.y, b.z}, {c.x, c.y, c.z}};
        for (auto &v : vs)
            for (int k = 0; k < 3; k++) l.b.lo[k] = std::fmin(l.b.lo[k], v[k]), l.b.hi[k] = std::fmax(l.b.hi[k], v[k]);
        l.b.cos_theta_o = std::cos(l.b.theta_o), l.b.cos_theta_e = std::cos(l.b.theta_e);
        lights.push_back(l);
    }
    if (lights.empty()) {
        err = "ats: no emissive triangle (LightSamplerATS::new(..).unwrap() panics)";
        return false;
    }
    Builder bd;
    bool ok = true;
    out.root = bd.build(lights.data(), lights.size(), 1, &ok);
    if (!ok) {
        err = "ats: a split left one side empty (build_bvh reaches unimplemented!() in the reference)";
        return false;
    }
    const size_t nn = bd.node_bounds.size();
    out.nodes.resize(4 * nn);
    out.leaf_of_prim.assign(ntris, 0xffffffffu);
    out.depth = 0;
    for (size_t i = 0; i < nn; i++) {
        const AtsBounds &b = bd.node_bounds[i];
        const float c[3] = {center_of(b, 0), center_of(b, 1), center_of(b, 2)};
        const float d[3] = {c[0] - b.hi[0], c[1] - b.hi[1], c[2] - b.hi[2]};
        out.nodes[4 * i] = make_float4(c[0], c[1], c[2], len3(d)); // AABB::to_sphere, structure.rs:871-877
        out.nodes[4 * i + 1] = make_float4(b.w[0], b.w[1], b.w[2], b.phi);
        out.nodes[4 * i + 2] = make_float4(b.cos_theta_o, b.cos_theta_e, b.two_sided ? 1.0f : 0.0f, 0.0f);
        out.nodes[4 * i + 3] = make_float4(u2f(bd.left[i]), u2f(bd.right[i]), u2f(bd.parent[i]), u2f(bd.light[i]));
        if (bd.light[i] != 0xffffffffu) out.leaf_of_prim[bd.light[i]] = (uint32_t)i;
        out.depth = out.depth > bd.depth_of[i] ? out.depth : bd.depth_of[i];
    }
    return true;
}

} // namespace rl
