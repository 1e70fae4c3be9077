// rl_refbvh_host.hpp -- the reference's own BVH, rebuilt on the host (once per scene) for ONE purpose: its visit order.
//
// The device finds the closest accepted triangle with its own conservative structures (quad table, LBVH); when two accepted
// hits lie within 2e-5 t of each other -- exact ties on shared edges and wall seams, coplanar faces -- the answer of the
// reference depends on the order in which BVHAccel visits its leaves (strict `t < its.t`, accel.rs:243-288 + geometry.rs:398)
// and on its box culling.  Such rays (a few in 10^4 of a pixel-centre grid, far fewer of jittered rays) are re-traced by
// rl_device.cuh: ref_bvh_closest over the tree built here, which restates
//   BVHAccel::new + subdivide_node   accel.rs:107-240   (SAH sweep over the three axes, leaf at <= 2 primitives)
//   Mesh::compute_aabb_tri           geometry.rs:423-439 (flat extents padded by +-1e-4)
//   AABB::{union_aabb, center, surface_area}   structure.rs:779-846 (surface_area is half the true area)
// The comparator of the reference's sort_by never returns Equal; Rust's sort_by is a stable merge sort that only asks
// `compare(a, b) == Less`, so it behaves as a stable sort on `<` (std::stable_sort below).  rustc's exact sort algorithm is
// not pinned by anything in the reference tree: for primitives with EQUAL box centres on the split axis this is an assumption.
#pragma once
#include <algorithm>
#include <cstdint>
#include <limits>
#include <vector>

#include "rl_scene_host.hpp"

namespace rl {

struct RefBVH {
    std::vector<float4> nodes;   // 2 per node: {p_min, info} {p_max, count}  (count 0 = inner node with children info, info + 1)
    std::vector<uint32_t> prims; // leaf contents: ORIGINAL triangle index (mesh-major), in the reference's primitive order
    std::vector<uint32_t> up;    // [node] parent of the node (root: 0), then [n_nodes + i] the leaf that holds prims[i]
    uint32_t depth = 0;
};

struct RefBox {
    float lo[3], hi[3];
    uint32_t prim;
};
inline void refbox_reset(RefBox &b) {
    for (int a = 0; a < 3; a++) b.lo[a] = std::numeric_limits<float>::max(), b.hi[a] = -std::numeric_limits<float>::max(); // f32::MAX / f32::MIN
}
inline void refbox_union(RefBox &acc, const RefBox &b) {
    for (int a = 0; a < 3; a++) acc.lo[a] = std::fmin(acc.lo[a], b.lo[a]), acc.hi[a] = std::fmax(acc.hi[a], b.hi[a]);
}
inline float refbox_center(const RefBox &b, int a) { return (b.hi[a] - b.lo[a]) * 0.5f + b.lo[a]; } // size() * 0.5 + p_min
inline float refbox_half_area(const RefBox &b) { // sum over i of the product of the two other extents, in the reference's order
    const float d[3] = {b.hi[0] - b.lo[0], b.hi[1] - b.lo[1], b.hi[2] - b.lo[2]};
    float s = 0.0f;
    for (int i = 0; i < 3; i++) {
        float v = 1.0f;
        for (int j = 0; j < 3; j++)
            if (j != i) v *= d[j];
        s += v;
    }
    return s;
}

inline void build_ref_bvh(const HostScene &hs, RefBVH &out) {
    out = RefBVH{};
    const uint32_t n = hs.ntris;
    std::vector<RefBox> boxes(n);
    RefBox root;
    refbox_reset(root);
    for (uint32_t p = 0; p < n; p++) { // compute_aabb_tri in mesh-major order
        RefBox &b = boxes[p];
        refbox_reset(b);
        b.prim = p;
        for (int k = 0; k < 3; k++) {
            const float4 v = hs.verts[3 * (size_t)p + k];
            const float c[3] = {v.x, v.y, v.z};
            for (int a = 0; a < 3; a++) b.lo[a] = std::fmin(b.lo[a], c[a]), b.hi[a] = std::fmax(b.hi[a], c[a]);
        }
        for (int a = 0; a < 3; a++)
            if (b.hi[a] - b.lo[a] < RL_EPSILON) b.hi[a] += RL_EPSILON, b.lo[a] -= RL_EPSILON;
        refbox_union(root, b);
    }
    struct Node {
        RefBox box;
        size_t info, count;
    };
    std::vector<Node> nodes;
    nodes.push_back(Node{root, 0, n});
    struct Work {
        size_t node;
        uint32_t depth;
    };
    std::vector<Work> todo{{0, 1}};
    std::vector<float> scores;
    while (!todo.empty()) { // depth first, left child before right: node indices come out in the order of the reference's recursion
        const Work w = todo.back();
        todo.pop_back();
        out.depth = std::max(out.depth, w.depth);
        if (nodes[w.node].count <= 2) continue;
        const size_t count = nodes[w.node].count, first = nodes[w.node].info;
        nodes[w.node].count = 0;
        nodes[w.node].info = nodes.size();
        auto sort_axis = [&](int axis) {
            std::stable_sort(boxes.begin() + first, boxes.begin() + first + count,
                             [axis](const RefBox &a, const RefBox &b) { return refbox_center(a, axis) < refbox_center(b, axis); });
        };
        size_t best_pos = 0;
        float best_cost = std::numeric_limits<float>::infinity();
        int best_axis = 3;
        scores.assign(count - 1, 0.0f);
        for (int axis = 0; axis < 3; axis++) {
            sort_axis(axis);
            RefBox acc;
            refbox_reset(acc);
            for (size_t id = 0; id + 1 < count; id++) { // right-to-left sweep
                const size_t id_left = count - id - 1;
                refbox_union(acc, boxes[first + id_left]);
                scores[id_left - 1] = refbox_half_area(acc) * (float)(id + 1);
            }
            refbox_reset(acc);
            for (size_t id = 0; id + 1 < count; id++) { // left-to-right sweep
                refbox_union(acc, boxes[first + id]);
                scores[id] += refbox_half_area(acc) * (float)(id + 1);
                if (scores[id] < best_cost) best_cost = scores[id], best_axis = axis, best_pos = id + 1;
            }
        }
        if (best_axis < 3) sort_axis(best_axis); // (axis 3 would index out of bounds in the reference: unreachable unless every score is NaN)
        size_t offset = best_pos;
        if (best_pos == count || best_pos == 0) offset = std::max<size_t>((size_t)((float)count * 0.5f), 1);
        Node left, right;
        refbox_reset(left.box), refbox_reset(right.box);
        for (size_t i = 0; i < offset; i++) refbox_union(left.box, boxes[first + i]);
        for (size_t i = offset; i < count; i++) refbox_union(right.box, boxes[first + i]);
        left.info = first, left.count = offset;
        right.info = first + offset, right.count = count - offset;
        const size_t id_left = nodes.size();
        nodes.push_back(left);
        nodes.push_back(right);
        todo.push_back(Work{id_left + 1, w.depth + 1});
        todo.push_back(Work{id_left, w.depth + 1});
    }
    out.nodes.resize(2 * nodes.size());
    for (size_t i = 0; i < nodes.size(); i++) {
        const Node &nd = nodes[i];
        out.nodes[2 * i] = f4(nd.box.lo[0], nd.box.lo[1], nd.box.lo[2], u2f((uint32_t)nd.info));
        out.nodes[2 * i + 1] = f4(nd.box.hi[0], nd.box.hi[1], nd.box.hi[2], u2f((uint32_t)nd.count));
    }
    out.prims.resize(n);
    for (uint32_t i = 0; i < n; i++) out.prims[i] = boxes[i].prim;
    out.up.assign(nodes.size() + n, 0u);
    for (size_t i = 0; i < nodes.size(); i++) {
        if (nodes[i].count == 0) out.up[nodes[i].info] = (uint32_t)i, out.up[nodes[i].info + 1] = (uint32_t)i;
        else
            for (size_t k = 0; k < nodes[i].count; k++) out.up[nodes.size() + nodes[i].info + k] = (uint32_t)i;
    }
}

} // namespace rl
