// rl_wide_host.hpp -- the tree the traversal kernels walk on scenes without a group table, built on the host from the TOPOLOGY of the
// reference's own BVH (rl_refbvh_host.hpp = BVHAccel::new, accel.rs:107-240: full SAH sweep, leaves of <= 2 primitives).
//
// Why the reference's topology: a sweep-SAH tree costs incoherent rays far fewer node visits than a Morton-order LBVH, the tree is
// needed on the host anyway (it decides ties and rim hits), and the reference pays the same build.  Why not its boxes: the device's
// slab test is fma-based and must stay conservative, so every child box is the reference's box grown by bvh_box_eps (rl_build.cuh).
// Culling that is merely conservative cannot change a result (rl_device.cuh: the exact triangle test + the tie machinery decide).
//
//   width 2: the 64-byte node of rl_build.cuh (write_wide_node): two child boxes + two child references.
//   width 4: inner nodes of the binary tree are collapsed (the child with the largest box is opened until the node has four
//            children or only leaves): seven float4 per node in SoA form
//              [0] lo.x of children 0-3   [1] lo.y   [2] lo.z   [3] hi.x   [4] hi.y   [5] hi.z   [6] child references (int bits)
//            One visit = seven 128-bit loads for four slab tests instead of two dependent visits of four loads each: half the
//            dependent memory round trips per ray on incoherent rays.  Unused child slots hold RL_TRAV_EMPTY and an inverted infinite box.
// Child references as in rl_device.cuh: >= 0 inner node index, bit 31 set = leaf_ref(first slot, count).  Slot order = the
// reference's primitive order (RefBVH::prims), so every leaf covers consecutive slots.
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

#include "rl_build.cuh"
#include "rl_refbvh_host.hpp"

namespace rl {

struct WideTree {
    std::vector<float4> nodes;
    uint32_t width = 2;      // 2 or 4
    uint32_t n_nodes = 0, n_leaves = 0;
    uint32_t depth = 0;      // inner levels + 1
    uint32_t max_stack = 0;  // bound on the traversal stack: (width - 1) entries per inner level
};

// `rb` as build_ref_bvh returns it (prims = original triangle indices in leaf order); needs more than 2 triangles (an inner root).
inline void build_wide_tree(const RefBVH &rb, float eps, uint32_t width, WideTree &out) {
    out = WideTree{};
    out.width = width;
    const size_t nn = rb.nodes.size() / 2;
    auto is_inner = [&](uint32_t i) { return f2u(rb.nodes[2 * i + 1].w) == 0u; };
    auto first_child = [&](uint32_t i) { return f2u(rb.nodes[2 * i].w); };
    auto half_area = [&](uint32_t i) {
        const float4 a = rb.nodes[2 * i], b = rb.nodes[2 * i + 1];
        const float dx = b.x - a.x, dy = b.y - a.y, dz = b.z - a.z;
        return dx * dy + dy * dz + dz * dx;
    };
    struct Item {
        uint32_t ref_node; // inner node of the binary tree that becomes wide node `index`
        uint32_t index, depth;
    };
    std::vector<Item> todo{{0u, 0u, 1u}};
    uint32_t n_wide = 1;
    const uint32_t f4_per = width == 4 ? 7u : 4u;
    out.nodes.assign(f4_per, f4(0, 0, 0, 0));
    while (!todo.empty()) {
        const Item it = todo.back();
        todo.pop_back();
        out.depth = out.depth > it.depth ? out.depth : it.depth;
        uint32_t kids[4];
        uint32_t nk = 2;
        kids[0] = first_child(it.ref_node), kids[1] = kids[0] + 1;
        while (nk < width) { // open the inner child with the largest box
            int best = -1;
            float best_a = -1.0f;
            for (uint32_t k = 0; k < nk; k++)
                if (is_inner(kids[k]) && half_area(kids[k]) > best_a) best_a = half_area(kids[k]), best = (int)k;
            if (best < 0) break;
            const uint32_t c = first_child(kids[best]);
            kids[best] = c;
            kids[nk++] = c + 1;
        }
        int refs[4];
        V3 lo[4], hi[4];
        for (uint32_t k = 0; k < 4; k++) {
            refs[k] = RL_TRAV_EMPTY;
            // an inverted, infinite box: the slab test of an unused slot fails by itself (entry +inf, exit -inf whatever the direction;
            // inf x (1/d clamped to a finite non-zero value) is never NaN)
            lo[k] = V3{INFINITY, INFINITY, INFINITY}, hi[k] = V3{-INFINITY, -INFINITY, -INFINITY};
        }
        for (uint32_t k = 0; k < nk; k++) {
            const float4 a = rb.nodes[2 * kids[k]], b = rb.nodes[2 * kids[k] + 1];
            lo[k] = V3{a.x - eps, a.y - eps, a.z - eps}, hi[k] = V3{b.x + eps, b.y + eps, b.z + eps};
            if (is_inner(kids[k])) {
                refs[k] = (int)n_wide;
                todo.push_back(Item{kids[k], n_wide, it.depth + 1});
                n_wide++;
                out.nodes.resize((size_t)n_wide * f4_per, f4(0, 0, 0, 0));
            } else {
                refs[k] = leaf_ref(f2u(a.w), f2u(b.w));
                out.n_leaves++;
            }
        }
        float4 *nd = out.nodes.data() + (size_t)it.index * f4_per;
        if (width == 4) {
            nd[0] = f4(lo[0].x, lo[1].x, lo[2].x, lo[3].x), nd[1] = f4(lo[0].y, lo[1].y, lo[2].y, lo[3].y), nd[2] = f4(lo[0].z, lo[1].z, lo[2].z, lo[3].z);
            nd[3] = f4(hi[0].x, hi[1].x, hi[2].x, hi[3].x), nd[4] = f4(hi[0].y, hi[1].y, hi[2].y, hi[3].y), nd[5] = f4(hi[0].z, hi[1].z, hi[2].z, hi[3].z);
            nd[6] = f4(u2f((uint32_t)refs[0]), u2f((uint32_t)refs[1]), u2f((uint32_t)refs[2]), u2f((uint32_t)refs[3]));
        } else {
            write_wide_node(out.nodes.data(), (int)it.index, lo[0], hi[0], lo[1], hi[1], refs[0], refs[1]);
        }
    }
    (void)nn;
    out.n_nodes = n_wide;
    out.max_stack = (width - 1u) * out.depth + 1u;
}

} // namespace rl
