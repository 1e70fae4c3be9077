// rl_flat_host.hpp -- builds the flat QUAD table scanned by rl_device.cuh: flat_scan (host side, once per
// scene, scenes of at most 64 triangles).  The table only feeds the conservative prefilter: every accepted
// hit still goes through the exact triangle test on the trav[] records, so nothing here can change a result,
// only how many exact tests run.
//
// Quad record: two triangles A, B share a record when their float normals agree to a few ulps (up to sign) and
// the vertices of B lie within 1e-6 * abs_max of the FLOAT plane of A (measured in double).  The largest such
// mismatch is returned as `delta`; flat_ray() adds it to the t margin (a point of B that the ray reaches at t_B
// is within delta of A's plane, so |t_A - t_B| <= delta / |d.n|).  Unpaired triangles get a record of their own.
//
// Per record the table holds the plane of A and ONE bounding parallelogram of A u B in that plane, as two affine
// functionals U(p), V(p) with 0 <= U, V <= 1 on every vertex (a pair of A's barycentric functionals u, v, w = 1-u-v,
// rescaled to the vertex ranges; the pair with the smallest parallelogram is taken, which for a quad split along a
// diagonal is the frame at the vertex opposite to it: area = area(A) + area(B) for a parallelogram).  When A and B
// share an edge bit for bit and lie on opposite sides of it, the record also carries the functional D = barycentric
// weight of A's vertex opposite to the shared edge, as D = k0 + kU U + kV V: D >= 0 on A, D <= 0 on B, so the scan can
// tell which of the two triangles a ray can hit.  Records that cannot be separated carry k = 0 (both are candidates).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "rl_build.cuh"
#include "rl_scene_host.hpp"

namespace rl {

struct FlatTable {
    std::vector<float4> f4;     // RL_FLAT_F4 per group + RL_FLAT_TAIL_F4 (slot bytes)
    uint32_t n_groups = 0;
    uint32_t valid_a = 0, valid_b = 0; // candidate bits of real triangles (A / B of each quad), bit order of flat_scan()
    float delta = 0.0f;
    uint32_t n_pairs = 0, n_singles = 0, n_separable = 0;
    // per quad bit: the vertices of A and B (world space, for the camera-tile culling kernel): 6 x float4 per bit, B = A when absent
    std::vector<float4> quad_verts;
};

struct D3 {
    double x, y, z;
};
inline D3 d3(float4 v) { return D3{(double)v.x, (double)v.y, (double)v.z}; }
inline D3 operator-(D3 a, D3 b) { return D3{a.x - b.x, a.y - b.y, a.z - b.z}; }
inline D3 operator+(D3 a, D3 b) { return D3{a.x + b.x, a.y + b.y, a.z + b.z}; }
inline D3 operator*(D3 a, double s) { return D3{a.x * s, a.y * s, a.z * s}; }
inline double ddot(D3 a, D3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline D3 dcross(D3 a, D3 b) { return D3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
struct Affine { // f(p) = m.p + c
    D3 m;
    double c;
    double at(D3 p) const { return ddot(m, p) + c; }
};

// prim_of_slot[s] = original triangle index at Morton slot s (the order of the trav[] records).
inline bool build_flat_table(const HostScene &hs, const std::vector<uint32_t> &prim_of_slot, FlatTable &out) {
    const uint32_t n = hs.ntris;
    out = FlatTable{};
    if (n == 0 || n > 64 || prim_of_slot.size() != n) return false;
    std::vector<float4> rec((size_t)RL_TRAV_F4 * n), shade_tmp(hs.shade);
    for (uint32_t s = 0; s < n; s++) tri_setup(hs.verts.data(), prim_of_slot[s], s, bvh_box_eps(hs.abs_max), rec.data(), shade_tmp.data());
    const double cap = 1e-6 * (double)hs.abs_max;
    auto vert = [&](uint32_t slot, int k) { return hs.verts[3 * (size_t)prim_of_slot[slot] + k]; };
    auto mismatch = [&](uint32_t a, uint32_t b) { // vertices of slot b against the float plane of slot a
        const float4 pl = rec[(size_t)RL_TRAV_F4 * a + 3];
        double worst = 0.0;
        for (int k = 0; k < 3; k++) {
            const float4 v = vert(b, k);
            double dist = std::fabs((double)pl.x * v.x + (double)pl.y * v.y + (double)pl.z * v.z - (double)pl.w);
            if (!(dist <= worst)) worst = dist; // NaN propagates as "large"
        }
        return worst;
    };
    auto normals_agree = [&](uint32_t a, uint32_t b) {
        const float4 na = rec[(size_t)RL_TRAV_F4 * a + 3], nb = rec[(size_t)RL_TRAV_F4 * b + 3];
        float same = fmaxf(fmaxf(fabsf(na.x - nb.x), fabsf(na.y - nb.y)), fabsf(na.z - nb.z));
        float opp = fmaxf(fmaxf(fabsf(na.x + nb.x), fabsf(na.y + nb.y)), fabsf(na.z + nb.z));
        return same <= 4e-7f || opp <= 4e-7f; // false for NaN normals (degenerate triangles stay single)
    };
    auto same_vertex = [&](float4 p, float4 q) { return std::memcmp(&p, &q, 12) == 0; };
    auto shares_edge = [&](uint32_t a, uint32_t b) { // two vertices of b equal two vertices of a bit for bit
        int cnt = 0;
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++)
                if (same_vertex(vert(a, i), vert(b, j))) {
                    cnt++;
                    break;
                }
        return cnt >= 2;
    };
    struct Pair {
        uint32_t a, b;
        bool has_b;
    };
    std::vector<Pair> pairs;
    std::vector<char> used(n, 0);
    double delta = 0.0;
    for (uint32_t a = 0; a < n; a++) {
        if (used[a]) continue;
        used[a] = 1;
        int best = -1;
        double best_d = cap;
        bool best_shares = false;
        for (uint32_t b = a + 1; b < n; b++) { // prefer a partner that shares an edge (a quad split along a diagonal), then the flattest
            if (used[b] || !normals_agree(a, b)) continue;
            double d = mismatch(a, b);
            if (!(d <= cap)) continue;
            const bool sh = shares_edge(a, b);
            if (best < 0 || (sh && !best_shares) || (sh == best_shares && d <= best_d)) best_d = d, best = (int)b, best_shares = sh;
        }
        if (best >= 0) {
            used[best] = 1;
            pairs.push_back(Pair{a, (uint32_t)best, true});
            if (best_d > delta) delta = best_d;
            out.n_pairs++;
        } else {
            pairs.push_back(Pair{a, a, false});
            out.n_singles++;
        }
    }
    const uint32_t n_groups = (uint32_t)((pairs.size() + 1) / 2);
    if (n_groups > RL_FLAT_MAX_GROUPS) return false;
    out.n_groups = n_groups;
    out.delta = (float)(delta * 1.0000002) + 1e-30f;
    out.f4.assign((size_t)RL_FLAT_F4 * n_groups + RL_FLAT_TAIL_F4, f4(0, 0, 0, 0));
    out.quad_verts.assign((size_t)6 * 32, f4(0, 0, 0, 0));
    unsigned char *slot_tab = reinterpret_cast<unsigned char *>(&out.f4[(size_t)RL_FLAT_F4 * n_groups]); // [bit] A slot, [32 + bit] B slot

    struct Quad { // one record, in float as the table stores it
        float n[4];          // plane of A: n.p = pn
        float mu[4], mv[4];  // U(p) = mu.xyz . p + mu.w
        float mn, k0, ku, kv;
    };
    auto make_quad = [&](const Pair &P, Quad &q) -> bool /* separable */ {
        const float4 pl = rec[(size_t)RL_TRAV_F4 * P.a + 3];
        q.n[0] = pl.x, q.n[1] = pl.y, q.n[2] = pl.z, q.n[3] = pl.w;
        // barycentric functionals of A in double, from the float vertices
        const D3 a0 = d3(vert(P.a, 0)), a1 = d3(vert(P.a, 1)), a2 = d3(vert(P.a, 2));
        const D3 e1 = a1 - a0, e2 = a2 - a0, cr = dcross(e1, e2);
        const double det = std::sqrt(ddot(cr, cr));
        const D3 nh = cr * (1.0 / det);
        Affine fu, fv, fw;
        fu.m = dcross(e2, nh) * (1.0 / det), fu.c = -ddot(a0, fu.m);
        fv.m = dcross(nh, e1) * (1.0 / det), fv.c = -ddot(a0, fv.m);
        fw.m = (fu.m + fv.m) * -1.0, fw.c = 1.0 - fu.c - fv.c;
        std::vector<D3> pts = {a0, a1, a2};
        if (P.has_b)
            for (int k = 0; k < 3; k++) pts.push_back(d3(vert(P.b, k)));
        const Affine *cand[3][2] = {{&fu, &fv}, {&fv, &fw}, {&fw, &fu}};
        double best_area = INFINITY;
        Affine U = fu, V = fv;
        for (auto &c : cand) {
            double lo[2] = {INFINITY, INFINITY}, hi[2] = {-INFINITY, -INFINITY};
            for (int j = 0; j < 2; j++)
                for (const D3 &p : pts) {
                    double v = c[j]->at(p);
                    lo[j] = std::fmin(lo[j], v), hi[j] = std::fmax(hi[j], v);
                }
            const double area = (hi[0] - lo[0]) * (hi[1] - lo[1]);
            if (area < best_area) { // degenerate triangles (NaN) never compare true: U, V stay NaN functionals => always candidates
                best_area = area;
                U.m = c[0]->m * (1.0 / (hi[0] - lo[0])), U.c = (c[0]->c - lo[0]) / (hi[0] - lo[0]);
                V.m = c[1]->m * (1.0 / (hi[1] - lo[1])), V.c = (c[1]->c - lo[1]) / (hi[1] - lo[1]);
            }
        }
        q.mu[0] = (float)U.m.x, q.mu[1] = (float)U.m.y, q.mu[2] = (float)U.m.z, q.mu[3] = (float)U.c;
        q.mv[0] = (float)V.m.x, q.mv[1] = (float)V.m.y, q.mv[2] = (float)V.m.z, q.mv[3] = (float)V.c;
        double gn = std::fmax(std::sqrt(ddot(U.m, U.m)), std::sqrt(ddot(V.m, V.m)));
        q.k0 = q.ku = q.kv = 0.0f;
        bool separable = false;
        if (P.has_b) {
            // shared edge: the vertex of A that is not (bit for bit) a vertex of B, when exactly one such vertex exists
            int lone = -1, n_lone = 0;
            for (int i = 0; i < 3; i++) {
                bool found = false;
                for (int j = 0; j < 3; j++) found = found || same_vertex(vert(P.a, i), vert(P.b, j));
                if (!found) lone = i, n_lone++;
            }
            int lone_b = -1, n_lone_b = 0;
            for (int j = 0; j < 3; j++) {
                bool found = false;
                for (int i = 0; i < 3; i++) found = found || same_vertex(vert(P.a, i), vert(P.b, j));
                if (!found) lone_b = j, n_lone_b++;
            }
            if (n_lone == 1 && n_lone_b == 1) {
                const Affine &D = lone == 0 ? fw : (lone == 1 ? fu : fv); // weight of A's lone vertex: 1 there, 0 on the shared edge
                const double db = D.at(d3(vert(P.b, lone_b)));
                if (db < -1e-3) { // B's lone vertex clearly on the other side: A = {D >= 0}, B = {D <= 0} inside the quad
                    // D = k0 + kU U + kV V on the plane: solve on A's three vertices
                    const double u0 = U.at(a0), v0 = V.at(a0), u1 = U.at(a1), v1 = V.at(a1), u2 = U.at(a2), v2 = V.at(a2);
                    const double dd0 = D.at(a0), dd1 = D.at(a1), dd2 = D.at(a2);
                    const double a11 = u1 - u0, a12 = v1 - v0, a21 = u2 - u0, a22 = v2 - v0, dt = a11 * a22 - a12 * a21;
                    if (std::fabs(dt) > 1e-12) {
                        const double ku = ((dd1 - dd0) * a22 - a12 * (dd2 - dd0)) / dt, kv = (a11 * (dd2 - dd0) - (dd1 - dd0) * a21) / dt;
                        q.ku = (float)ku, q.kv = (float)kv, q.k0 = (float)(dd0 - ku * u0 - kv * v0);
                        gn = std::fmax(gn, std::sqrt(ddot(D.m, D.m)));
                        separable = std::isfinite(q.ku) && std::isfinite(q.kv) && std::isfinite(q.k0);
                        if (!separable) q.k0 = q.ku = q.kv = 0.0f;
                    }
                }
            }
        }
        q.mn = (float)(gn * 1.000001);
        if (!(q.mn == q.mn)) q.mn = 0.0f; // NaN gradient (degenerate A): the functionals are NaN and keep the record a candidate
        return separable;
    };
    const Pair none{0, 0, false};
    for (uint32_t gi = 0; gi < n_groups; gi++) {
        const bool has_q = 2 * gi + 1 < pairs.size();
        const Pair P = pairs[2 * gi], Q = has_q ? pairs[2 * gi + 1] : P;
        Quad qp, qq;
        const bool sep_p = make_quad(P, qp), sep_q = make_quad(Q, qq);
        out.n_separable += (sep_p ? 1u : 0u) + ((has_q && sep_q) ? 1u : 0u);
        float4 *g = &out.f4[(size_t)RL_FLAT_F4 * gi];
        g[0] = f4(qp.n[0], qq.n[0], qp.n[1], qq.n[1]);
        g[1] = f4(qp.n[2], qq.n[2], qp.n[3], qq.n[3]);
        g[2] = f4(qp.mu[0], qq.mu[0], qp.mu[1], qq.mu[1]);
        g[3] = f4(qp.mu[2], qq.mu[2], qp.mu[3], qq.mu[3]);
        g[4] = f4(qp.mv[0], qq.mv[0], qp.mv[1], qq.mv[1]);
        g[5] = f4(qp.mv[2], qq.mv[2], qp.mv[3], qq.mv[3]);
        g[6] = f4(qp.mn, qq.mn, qp.k0, qq.k0);
        g[7] = f4(qp.ku, qq.ku, qp.kv, qq.kv);
        // bit of quad j (scan index j = 2 gi + {0, 1}): 2 n_groups - 1 - j
        const uint32_t bit_p = 2u * n_groups - 1u - 2u * gi, bit_q = bit_p - 1u;
        auto fill = [&](uint32_t bit, const Pair &R) {
            out.valid_a |= 1u << bit;
            slot_tab[bit] = (unsigned char)R.a;
            slot_tab[32 + bit] = (unsigned char)R.b;
            if (R.has_b) out.valid_b |= 1u << bit;
            for (int k = 0; k < 3; k++) {
                out.quad_verts[6 * (size_t)bit + k] = vert(R.a, k);
                out.quad_verts[6 * (size_t)bit + 3 + k] = vert(R.b, k);
            }
        };
        fill(bit_p, P);
        if (has_q) fill(bit_q, Q);
    }
    (void)none;
    return true;
}

} // namespace rl
