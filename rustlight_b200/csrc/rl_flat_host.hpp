// rl_flat_host.hpp -- builds the flat group table scanned by rl_device.cuh: flat_scan (host side, once per
// scene, scenes of at most 64 triangles).  The table only feeds the conservative prefilter: every accepted
// hit still goes through the exact triangle test on the trav[] records, so nothing here can change a result,
// only how many exact tests run.
//
// Pairing: two triangles share a pair record when their float normals agree to a few ulps (up to sign) and
// the vertices of the second lie within 1e-6 * abs_max of the FLOAT plane of the first (measured in double).
// The largest such mismatch is returned as `delta`; flat_ray() adds it to the t margin (a point of B that
// the ray reaches at t_B is within delta of A's plane, so |t_A - t_B| <= delta / |d.n|).
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

#include "rl_build.cuh"
#include "rl_scene_host.hpp"

namespace rl {

struct FlatTable {
    std::vector<float4> f4;     // RL_FLAT_F4 per group
    uint32_t n_groups = 0;
    uint32_t valid[2] = {0, 0}; // candidate bits of real triangles, bit order of flat_candidates()
    float delta = 0.0f;
    uint32_t n_pairs = 0, n_singles = 0;
};

// prim_of_slot[s] = original triangle index at Morton slot s (the order of the trav[] records).
inline bool build_flat_table(const HostScene &hs, const std::vector<uint32_t> &prim_of_slot, FlatTable &out) {
    const uint32_t n = hs.ntris;
    out = FlatTable{};
    if (n == 0 || n > 64 || prim_of_slot.size() != n) return false;
    std::vector<float4> rec((size_t)RL_TRAV_F4 * n), shade_tmp(hs.shade);
    for (uint32_t s = 0; s < n; s++) tri_setup(hs.verts.data(), prim_of_slot[s], s, rec.data(), shade_tmp.data());
    const double cap = 1e-6 * (double)hs.abs_max;
    auto mismatch = [&](uint32_t a, uint32_t b) { // vertices of slot b against the float plane of slot a
        const float4 pl = rec[(size_t)RL_TRAV_F4 * a + 3];
        double worst = 0.0;
        for (int k = 0; k < 3; k++) {
            const float4 v = hs.verts[3 * (size_t)prim_of_slot[b] + k];
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
        for (uint32_t b = a + 1; b < n; b++) {
            if (used[b] || !normals_agree(a, b)) continue;
            double d = mismatch(a, b);
            if (d <= best_d) best_d = d, best = (int)b;
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
    out.f4.assign((size_t)RL_FLAT_F4 * n_groups, f4(0, 0, 0, 0));
    for (uint32_t gi = 0; gi < n_groups; gi++) {
        const bool has_q = 2 * gi + 1 < pairs.size();
        const Pair P = pairs[2 * gi], Q = has_q ? pairs[2 * gi + 1] : Pair{P.a, P.a, false};
        float4 *g = &out.f4[(size_t)RL_FLAT_F4 * gi];
        auto R = [&](uint32_t slot, int k) { return rec[(size_t)RL_TRAV_F4 * slot + k]; };
        auto weave = [&](float4 p, float4 q, float4 *lo, float4 *hi) { // {x P,Q  y P,Q} {z P,Q  w P,Q}
            *lo = f4(p.x, q.x, p.y, q.y);
            *hi = f4(p.z, q.z, p.w, q.w);
        };
        weave(R(P.a, 3), R(Q.a, 3), &g[0], &g[1]);
        weave(R(P.a, 4), R(Q.a, 4), &g[2], &g[3]);
        weave(R(P.a, 5), R(Q.a, 5), &g[4], &g[5]);
        weave(R(P.b, 4), R(Q.b, 4), &g[6], &g[7]);
        weave(R(P.b, 5), R(Q.b, 5), &g[8], &g[9]);
        const float mnP = fmaxf(R(P.a, 2).w, R(P.b, 2).w), mnQ = fmaxf(R(Q.a, 2).w, R(Q.b, 2).w);
        g[10] = f4(mnP, mnQ, u2f(P.a | (P.b << 8)), u2f(Q.a | (Q.b << 8)));
        // valid bits: scan index inside the half -> bit 4*(groups in half)-1-idx
        const int half = gi >= 8 ? 1 : 0;
        const uint32_t g0 = half ? 8u : 0u, g1 = half ? n_groups : (n_groups < 8u ? n_groups : 8u);
        const uint32_t nb = 4u * (g1 - g0), base = 4u * (gi - g0);
        const bool v[4] = {true, P.has_b, has_q, has_q && Q.has_b};
        for (uint32_t j = 0; j < 4; j++)
            if (v[j]) out.valid[half] |= 1u << (nb - 1u - (base + j));
    }
    return true;
}

} // namespace rl
