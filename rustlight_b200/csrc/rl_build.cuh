// rl_build.cuh -- per-element steps of the device LBVH build (replaces BVHAccel::new,
// src/accel.rs:201-240, whose recursive SAH sweep is sequential).  Morton codes + Karras'
// 2012 parallel radix-tree construction, one leaf per triangle, then a bottom-up box fit that
// stores both child boxes in the parent ("wide" node, 64 B = four 128-bit loads per visit).
// Tree shape does not influence results: culling is conservative and ties are broken by
// triangle index (rl_device.cuh), so hits equal the reference's brute-force semantics.
//
// Like rl_device.cuh these are RL_HD functions so the CPU emulator can run the same code.
#pragma once
#include "rl_device.cuh"

namespace rl {

RL_HD uint32_t expand_bits10(uint32_t v) { // 10 bits -> every third bit
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
RL_HD uint32_t clamp1023(float x) {
    float y = fminf(fmaxf(x * 1024.0f, 0.0f), 1023.0f);
    return (uint32_t)y;
}
// 64-bit key: 30-bit Morton code of the triangle's box centre, then the triangle index, so
// that keys are unique and the tree is fully determined.
RL_HD uint64_t morton_key(V3 lo, V3 hi, V3 smin, V3 sinv, uint32_t prim) {
    V3 c = V3{(lo.x + hi.x) * 0.5f, (lo.y + hi.y) * 0.5f, (lo.z + hi.z) * 0.5f};
    uint32_t x = clamp1023((c.x - smin.x) * sinv.x), y = clamp1023((c.y - smin.y) * sinv.y), z = clamp1023((c.z - smin.z) * sinv.z);
    uint32_t code = (expand_bits10(x) << 2) | (expand_bits10(y) << 1) | expand_bits10(z);
    return ((uint64_t)code << 32) | (uint64_t)prim;
}

RL_HD void tri_bounds(const float4 *verts, uint32_t prim, V3 *lo, V3 *hi) {
    V3 a = xyz(verts[3 * prim]), b = xyz(verts[3 * prim + 1]), c = xyz(verts[3 * prim + 2]);
    *lo = V3{fminf(fminf(a.x, b.x), c.x), fminf(fminf(a.y, b.y), c.y), fminf(fminf(a.z, b.z), c.z)};
    *hi = V3{fmaxf(fmaxf(a.x, b.x), c.x), fmaxf(fmaxf(a.y, b.y), c.y), fmaxf(fmaxf(a.z, b.z), c.z)};
}

// Leaf box used by the LBVH: the raw bounds grown by `eps` (3e-5 * abs_max) so that the
// fma-based slab test of rl_device.cuh stays conservative.
RL_HD float bvh_box_eps(float abs_max) { return 3e-5f * abs_max + 1e-30f; }
RL_HD void tri_bounds_inflated(const float4 *verts, uint32_t prim, float eps, V3 *lo, V3 *hi) {
    tri_bounds(verts, prim, lo, hi);
    *lo = V3{lo->x - eps, lo->y - eps, lo->z - eps};
    *hi = V3{hi->x + eps, hi->y + eps, hi->z + eps};
}

// Ray-independent part of Mesh::intersection_tri (geometry.rs:365-372, 383): e1, e2,
// n_geo = normalize(e1 x e2), det = |e1 x e2| with the reference's operations, plus the data of
// the conservative prefilter (rl_device.cuh: tri_prefilter): plane offset pn = v0.n and the
// affine barycentric functionals u(p) = Mu.p + cu, v(p) = Mv.p + cv.  Six float4 per triangle,
// written to Morton slot s; n_geo also goes to the shading table.
#define RL_TRAV_F4 6
// `box_eps` = bvh_box_eps(abs_max): a triangle whose box is thin (< 30 box_eps ~ 1e-3 of the scene, or padded) along two axes is
// flagged RL_PRIM_NEEDLE in the prim word: the reference's slab test can fail on such a box even for interior hits (rl_device.cuh: hit_unsafe).
RL_HD void tri_setup(const float4 *verts, uint32_t prim, uint32_t s, float box_eps, float4 *trav, float4 *shade) {
    V3 v0 = xyz(verts[3 * prim]), v1 = xyz(verts[3 * prim + 1]), v2 = xyz(verts[3 * prim + 2]);
    const float ext[3] = {fmaxf(fmaxf(v0.x, v1.x), v2.x) - fminf(fminf(v0.x, v1.x), v2.x), fmaxf(fmaxf(v0.y, v1.y), v2.y) - fminf(fminf(v0.y, v1.y), v2.y),
                          fmaxf(fmaxf(v0.z, v1.z), v2.z) - fminf(fminf(v0.z, v1.z), v2.z)};
    int thin = 0;
    for (int a = 0; a < 3; a++)
        if (!(ext[a] >= fmaxf(RL_EPSILON, 30.0f * box_eps))) thin++;
    const uint32_t prim_word = prim | (thin >= 2 ? RL_PRIM_NEEDLE : 0u);
    V3 e1 = v1 - v0, e2 = v2 - v0;
    V3 cr = cross(e1, e2);
    V3 n_geo = normalize(cr);
    float det = magnitude(cr);
    float idet = 1.0f / det;
    V3 mu = cross(e2, n_geo) * idet, mv = cross(n_geo, e1) * idet;
    float mn = fmaxf(magnitude(mu), magnitude(mv));
    trav[RL_TRAV_F4 * s + 0] = make_float4(v0.x, v0.y, v0.z, det);
    trav[RL_TRAV_F4 * s + 1] = make_float4(e1.x, e1.y, e1.z, u2f(prim_word));
    trav[RL_TRAV_F4 * s + 2] = make_float4(e2.x, e2.y, e2.z, mn);
    trav[RL_TRAV_F4 * s + 3] = make_float4(n_geo.x, n_geo.y, n_geo.z, dot(v0, n_geo));
    trav[RL_TRAV_F4 * s + 4] = make_float4(mu.x, mu.y, mu.z, -dot(v0, mu));
    trav[RL_TRAV_F4 * s + 5] = make_float4(mv.x, mv.y, mv.z, -dot(v0, mv));
    float4 s0 = shade[4 * prim];
    shade[4 * prim] = make_float4(n_geo.x, n_geo.y, n_geo.z, s0.w);
}

RL_HD int clz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __clzll((long long)x);
#else
    return x ? __builtin_clzll(x) : 64;
#endif
}
// Karras 2012, "Maximizing Parallelism in the Construction of BVHs, Octrees, and k-d Trees".
RL_HD int karras_delta(const uint64_t *keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    return clz64(keys[i] ^ keys[j]);
}
// Internal node i of n-1: children as node refs (>=0 internal, ~leaf for leaves).
RL_HD void karras_node(const uint64_t *keys, int n, int i, int *left, int *right, int *first, int *last) {
    int d = (karras_delta(keys, n, i, i + 1) - karras_delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    int dmin = karras_delta(keys, n, i, i - d);
    int lmax = 2;
    while (karras_delta(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int t = lmax / 2; t >= 1; t /= 2)
        if (karras_delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    int dnode = karras_delta(keys, n, i, j);
    int s = 0;
    int t = l;
    do {
        t = (t + 1) >> 1;
        if (karras_delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    int gamma = i + s * d + (d < 0 ? -1 : 0);
    int lo = i < j ? i : j, hi = i < j ? j : i;
    *first = lo;
    *last = hi;
    *left = (lo == gamma) ? ~gamma : gamma;
    *right = (hi == gamma + 1) ? ~(gamma + 1) : (gamma + 1);
}

// Child reference of a wide node: >= 0 inner node index; bit 31 set = leaf covering `count`
// consecutive triangles of the Morton order starting at `first` (subtrees with at most
// leaf_max triangles are collapsed into one leaf).
RL_HD int make_child_ref(int child, const int2v *ranges, int leaf_max) {
    if (child < 0) return leaf_ref((uint32_t)~child, 1u);
    int size = ranges[child].y - ranges[child].x + 1;
    return size <= leaf_max ? leaf_ref((uint32_t)ranges[child].x, (uint32_t)size) : child;
}
// Write the wide node: child boxes + child refs.
RL_HD void write_wide_node(float4 *nodes, int i, V3 lo0, V3 hi0, V3 lo1, V3 hi1, int c0, int c1) {
    nodes[4 * i + 0] = make_float4(lo0.x, lo0.y, lo0.z, hi0.x);
    nodes[4 * i + 1] = make_float4(hi0.y, hi0.z, lo1.x, lo1.y);
    nodes[4 * i + 2] = make_float4(lo1.z, hi1.x, hi1.y, hi1.z);
    nodes[4 * i + 3] = make_float4(u2f((uint32_t)c0), u2f((uint32_t)c1), 0.0f, 0.0f);
}

} // namespace rl
