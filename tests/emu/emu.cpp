// emu.cpp -- serial CPU emulator of the device arithmetic (TEST INFRASTRUCTURE).
//
// Compiles rustlight_b200/csrc/rl_device.cuh, rl_build.cuh and rl_scene_host.hpp with g++
// (RL_HD expands to `inline`) and drives them in the same order as the kernels of
// rl_kernels.cuh: Morton keys -> sort -> triangle records -> Karras tree -> box fit, then per
// path: raygen -> trace_closest -> path_step -> trace_visible, radiance accumulated per path
// and summed per pixel in sample order.  It lets `pytest -m "not gpu"` verify on a CPU box
// that the device code takes the same discrete decisions as the oracle.  It is never built
// into or loaded by the product library.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "rl_b200.h"
#include "rl_build.cuh"
#include "rl_device.cuh"
#include "rl_flat_host.hpp"
#include "rl_refbvh_host.hpp"
#include "rl_scene_host.hpp"
#include "rl_wide_host.hpp"

using namespace rl;

struct emu_scene {
    HostScene hs;
    std::vector<float4> trav, nodes;
    FlatTable flat;
    RefBVH ref;
    SceneView sv{};
    uint32_t max_depth = 0;
    int root_ref = 0, leaf_max = 1;
};

extern "C" {

emu_scene *emu_scene_create(const rl_scene_desc *desc, char *err, size_t errlen) {
    auto *s = new emu_scene;
    std::string e;
    if (!build_host_scene(desc, s->hs, e)) {
        if (err && errlen) {
            std::strncpy(err, e.c_str(), errlen - 1);
            err[errlen - 1] = 0;
        }
        delete s;
        return nullptr;
    }
    HostScene &hs = s->hs;
    const int n = (int)hs.ntris;
    // k_morton
    V3 smin = V3{hs.raw_min[0], hs.raw_min[1], hs.raw_min[2]};
    V3 ext = V3{hs.raw_max[0] - hs.raw_min[0], hs.raw_max[1] - hs.raw_min[1], hs.raw_max[2] - hs.raw_min[2]};
    V3 sinv = V3{ext.x > 0.0f ? 1.0f / ext.x : 0.0f, ext.y > 0.0f ? 1.0f / ext.y : 0.0f, ext.z > 0.0f ? 1.0f / ext.z : 0.0f};
    std::vector<uint64_t> keys(n);
    for (int p = 0; p < n; p++) {
        V3 lo, hi;
        tri_bounds(hs.verts.data(), p, &lo, &hi);
        keys[p] = morton_key(lo, hi, smin, sinv, (uint32_t)p);
    }
    std::sort(keys.begin(), keys.end()); // cub::DeviceRadixSort on the device
    // Traversal mode (the device uses: group table for small scenes, the 4-wide tree over the reference's topology otherwise; LBVH as
    // the fallback; the single big leaf is the pre-group-table flat path, kept as a cross-check):
    // RL_EMU_ACCEL = flat | leaf | tree (Morton LBVH) | sah2 | sah4 (rl_wide_host.hpp)
    std::string mode = "flat";
    if (const char *e = getenv("RL_EMU_ACCEL")) mode = e;
    RefBVH rb_tree;
    const bool use_sah = (mode == "sah2" || mode == "sah4") && n > 2;
    if (use_sah) { // slot order = the reference's primitive order
        build_ref_bvh(hs, rb_tree);
        for (int i = 0; i < n; i++) keys[i] = ((uint64_t)i << 32) | (uint64_t)rb_tree.prims[i];
    }
    // k_tri_setup
    s->trav.resize((size_t)RL_TRAV_F4 * n);
    std::vector<V3> leaf_lo(n), leaf_hi(n);
    for (int i = 0; i < n; i++) {
        uint32_t prim = (uint32_t)(keys[i] & 0xffffffffull);
        tri_setup(hs.verts.data(), prim, i, bvh_box_eps(hs.abs_max), s->trav.data(), hs.shade.data());
        tri_bounds_inflated(hs.verts.data(), prim, bvh_box_eps(hs.abs_max), &leaf_lo[i], &leaf_hi[i]);
    }
    if (mode == "flat" && n <= RL_LEAF_MAX_CAP) {
        std::vector<uint32_t> prim_of_slot(n);
        for (int i = 0; i < n; i++) prim_of_slot[i] = (uint32_t)(keys[i] & 0xffffffffull);
        if (!build_flat_table(hs, prim_of_slot, s->flat)) s->flat = FlatTable{};
    }
    // k_karras + k_fit
    int leaf_max = (n <= RL_LEAF_MAX_CAP && mode == "leaf") ? RL_LEAF_MAX_CAP : 2;
    if (const char *e = getenv("RL_LEAF_MAX")) leaf_max = std::max(1, std::min(RL_LEAF_MAX_CAP, atoi(e)));
    s->leaf_max = leaf_max;
    int n_nodes = n > 1 ? n - 1 : 1;
    s->nodes.resize((size_t)4 * n_nodes);
    s->root_ref = n <= leaf_max ? leaf_ref(0u, (uint32_t)n) : 0;
    s->max_depth = 1;
    if (use_sah) {
        WideTree wt;
        build_wide_tree(rb_tree, bvh_box_eps(hs.abs_max), mode == "sah4" ? 4u : 2u, wt);
        if (wt.max_stack > (uint32_t)RL_STACK_SIZE) build_wide_tree(rb_tree, bvh_box_eps(hs.abs_max), 2u, wt);
        s->nodes = wt.nodes;
        s->root_ref = 0;
        s->sv.wide4 = wt.width == 4u ? 1u : 0u;
        s->max_depth = wt.depth;
    } else if (n > 1) {
        std::vector<int> cl(n - 1), cr(n - 1);
        std::vector<int2v> ranges(n - 1);
        for (int i = 0; i < n - 1; i++) karras_node(keys.data(), n, i, &cl[i], &cr[i], &ranges[i].x, &ranges[i].y);
        std::vector<V3> nlo(n - 1), nhi(n - 1);
        struct Rec {
            static void fit(int node, const std::vector<int> &cl, const std::vector<int> &cr, const std::vector<int2v> &ranges, int leaf_max,
                            const std::vector<V3> &llo, const std::vector<V3> &lhi, std::vector<V3> &nlo, std::vector<V3> &nhi, float4 *nodes) {
                int a = cl[node], b = cr[node];
                if (a >= 0) fit(a, cl, cr, ranges, leaf_max, llo, lhi, nlo, nhi, nodes);
                if (b >= 0) fit(b, cl, cr, ranges, leaf_max, llo, lhi, nlo, nhi, nodes);
                V3 lo0 = a < 0 ? llo[~a] : nlo[a], hi0 = a < 0 ? lhi[~a] : nhi[a];
                V3 lo1 = b < 0 ? llo[~b] : nlo[b], hi1 = b < 0 ? lhi[~b] : nhi[b];
                write_wide_node(nodes, node, lo0, hi0, lo1, hi1, make_child_ref(a, ranges.data(), leaf_max), make_child_ref(b, ranges.data(), leaf_max));
                nlo[node] = V3{fminf(lo0.x, lo1.x), fminf(lo0.y, lo1.y), fminf(lo0.z, lo1.z)};
                nhi[node] = V3{fmaxf(hi0.x, hi1.x), fmaxf(hi0.y, hi1.y), fmaxf(hi0.z, hi1.z)};
            }
        };
        Rec::fit(0, cl, cr, ranges, leaf_max, leaf_lo, leaf_hi, nlo, nhi, s->nodes.data());
    }
    // reference-order tree (tie rule), leaf contents as Morton slots
    if (!getenv("RL_NO_REF_ORDER")) {
        build_ref_bvh(hs, s->ref);
        if (s->ref.depth + 2 <= (uint32_t)RL_STACK_SIZE) {
            std::vector<uint32_t> slot_of_prim(n);
            for (int i = 0; i < n; i++) slot_of_prim[(uint32_t)(keys[i] & 0xffffffffull)] = (uint32_t)i;
            for (auto &p : s->ref.prims) p = slot_of_prim[p];
            const size_t nn = s->ref.nodes.size() / 2;
            std::vector<uint32_t> leaf_of_slot(n);
            for (int i = 0; i < n; i++) leaf_of_slot[s->ref.prims[i]] = s->ref.up[nn + i];
            for (int i = 0; i < n; i++) s->ref.up[nn + i] = leaf_of_slot[i];
        } else s->ref = RefBVH{};
    }
    SceneView &sv = s->sv;
    sv.ref_nodes = s->ref.nodes.empty() ? nullptr : s->ref.nodes.data(), sv.ref_prims = s->ref.prims.empty() ? nullptr : s->ref.prims.data();
    sv.ref_up = s->ref.up.empty() ? nullptr : s->ref.up.data(), sv.ref_n_nodes = (uint32_t)(s->ref.nodes.size() / 2);
    sv.trav = s->trav.data(), sv.nodes = s->nodes.data(), sv.shade = hs.shade.data(), sv.verts = hs.verts.data(), sv.mats = hs.mats.data();
    sv.emit_info = hs.emit_info.data(), sv.emit_cdf = hs.emit_cdf.data(), sv.area_cdf = hs.area_cdf.data();
    sv.uvs = hs.uvs.empty() ? nullptr : hs.uvs.data(), sv.tex = hs.tex.empty() ? nullptr : hs.tex.data(), sv.texels = hs.texels.data();
    sv.emit_var = hs.emit_var ? 1u : 0u;
    sv.ntris = hs.ntris, sv.n_emitters = hs.n_emitters;
    sv.root_ref = s->root_ref;
    sv.flat = s->flat.f4.data(), sv.n_groups = s->flat.n_groups, sv.flat_valid_a = s->flat.valid_a, sv.flat_valid_b = s->flat.valid_b;
    sv.flat_delta = s->flat.delta;
    sv.root_min = V3{hs.root_min[0], hs.root_min[1], hs.root_min[2]};
    sv.root_max = V3{hs.root_max[0], hs.root_max[1], hs.root_max[2]};
    sv.abs_max = hs.abs_max;
    sv.env_on = hs.env_on ? 1u : 0u, sv.env_color = Col{hs.env_color[0], hs.env_color[1], hs.env_color[2]};
    sv.bs_center = V3{hs.bs_center[0], hs.bs_center[1], hs.bs_center[2]}, sv.bs_radius = hs.bs_radius, sv.env_pdf_sel = hs.env_pdf_sel;
    sv.env_w = hs.env_w, sv.env_h = hs.env_h, sv.env_texel_off = hs.env_texel_off, sv.env_dist = hs.env_dist.data(), sv.env_func_int = hs.env_func_int;
    sv.ats_nodes = hs.ats_nodes.empty() ? nullptr : hs.ats_nodes.data(), sv.ats_leaf_of_prim = hs.ats_leaf_of_prim.data(), sv.ats_root = hs.ats_root;
    std::memcpy(sv.s2c, hs.s2c, 64);
    std::memcpy(sv.c2w, hs.c2w, 64);
    sv.cam_pos = V3{hs.cam_pos[0], hs.cam_pos[1], hs.cam_pos[2]};
    sv.img_w = (float)hs.img_w, sv.img_h = (float)hs.img_h;
    return s;
}
void emu_scene_destroy(emu_scene *s) { delete s; }
uint32_t emu_bvh_max_depth(const emu_scene *s) { return s->max_depth; }
// group table statistics: groups, pairs, singles; delta in *delta
uint32_t emu_flat_info(const emu_scene *s, uint32_t *pairs, uint32_t *singles, float *delta) {
    if (pairs) *pairs = s->flat.n_pairs;
    if (singles) *singles = s->flat.n_singles;
    if (delta) *delta = s->flat.delta;
    return s->flat.n_groups;
}

// Checks the tree: every triangle is referenced by exactly one leaf and every child box
// contains its subtree.  Returns 0 when valid.
int emu_bvh_validate(const emu_scene *s) {
    const int n = (int)s->hs.ntris;
    std::vector<int> seen(n, 0);
    struct Rec {
        static bool leaf(const emu_scene *s, int ref, V3 *lo, V3 *hi, std::vector<int> &seen) {
            *lo = V3{RL_F32_MAX, RL_F32_MAX, RL_F32_MAX};
            *hi = V3{-RL_F32_MAX, -RL_F32_MAX, -RL_F32_MAX};
            for (uint32_t k = 0; k < leaf_count(ref); k++) {
                uint32_t slot = leaf_first(ref) + k;
                if (slot >= s->hs.ntris) return false;
                seen[slot]++;
                V3 a, b;
                tri_bounds(s->hs.verts.data(), f2u(s->trav[RL_TRAV_F4 * slot + 1].w), &a, &b);
                *lo = V3{fminf(lo->x, a.x), fminf(lo->y, a.y), fminf(lo->z, a.z)};
                *hi = V3{fmaxf(hi->x, b.x), fmaxf(hi->y, b.y), fmaxf(hi->z, b.z)};
            }
            return true;
        }
        static bool walk(const emu_scene *s, int node, V3 *lo, V3 *hi, std::vector<int> &seen, uint32_t depth, uint32_t *max_depth) {
            *max_depth = std::max(*max_depth, depth);
            int ch[4], nq = 2;
            V3 blo[4], bhi[4];
            if (s->sv.wide4) {
                const float4 *nd = &s->nodes[7 * node];
                const float *f = &nd[0].x;
                nq = 0;
                for (int q = 0; q < 4; q++) {
                    const int c = (int)f2u(f[24 + q]);
                    if (c == RL_TRAV_EMPTY) continue;
                    ch[nq] = c, blo[nq] = V3{f[q], f[4 + q], f[8 + q]}, bhi[nq] = V3{f[12 + q], f[16 + q], f[20 + q]};
                    nq++;
                }
            } else {
                const float4 *nd = &s->nodes[4 * node];
                ch[0] = (int)f2u(nd[3].x), ch[1] = (int)f2u(nd[3].y);
                blo[0] = V3{nd[0].x, nd[0].y, nd[0].z}, blo[1] = V3{nd[1].z, nd[1].w, nd[2].x};
                bhi[0] = V3{nd[0].w, nd[1].x, nd[1].y}, bhi[1] = V3{nd[2].y, nd[2].z, nd[2].w};
            }
            for (int q = 0; q < nq; q++) {
                V3 clo, chi;
                if (ch[q] < 0) {
                    if (!leaf(s, ch[q], &clo, &chi, seen)) return false;
                } else if (!walk(s, ch[q], &clo, &chi, seen, depth + 1, max_depth)) return false;
                if (clo.x < blo[q].x || clo.y < blo[q].y || clo.z < blo[q].z || chi.x > bhi[q].x || chi.y > bhi[q].y || chi.z > bhi[q].z) return false;
            }
            *lo = blo[0], *hi = bhi[0];
            for (int q = 1; q < nq; q++) {
                *lo = V3{fminf(lo->x, blo[q].x), fminf(lo->y, blo[q].y), fminf(lo->z, blo[q].z)};
                *hi = V3{fmaxf(hi->x, bhi[q].x), fmaxf(hi->y, bhi[q].y), fmaxf(hi->z, bhi[q].z)};
            }
            return true;
        }
    };
    V3 lo, hi;
    uint32_t md = 1;
    bool ok = s->root_ref < 0 ? Rec::leaf(s, s->root_ref, &lo, &hi, seen) : Rec::walk(s, s->root_ref, &lo, &hi, seen, 1, &md);
    const_cast<emu_scene *>(s)->max_depth = md;
    if (!ok) return 1;
    for (int i = 0; i < n; i++)
        if (seen[i] != 1) return 2;
    return 0;
}

// Camera ray of pixel (px, py) the way k_trace_flat traces it on group-table scenes: only the quads the pixel's frustum overlaps
// (camera_block_mask; here a block of ONE pixel, the tightest mask there is -- the device uses blocks of 32 pixels).
static HitRec trace_camera(const emu_scene *s, uint32_t px, uint32_t py, V3 o, V3 d) {
    const SceneView &sv = s->sv;
    if (!sv.n_groups || getenv("RL_NO_CAM_CULL")) return trace_closest(sv, sv.nodes, sv.trav, o, d);
    const HostScene &hs = s->hs;
    const double ext = (double)hs.abs_max + std::max(std::max(std::fabs((double)hs.cam_pos[0]), std::fabs((double)hs.cam_pos[1])), std::fabs((double)hs.cam_pos[2]));
    const uint32_t quads = camera_block_mask(sv, s->flat.quad_verts.data(), s->flat.valid_a, px, px, py, py, 1e-4 * ext);
    HitRec h;
    h.t = RL_F32_MAX, h.u = 0.0f, h.v = 0.0f, h.prim = RL_MISS;
    V3 inv = V3{1.0f / d.x, 1.0f / d.y, 1.0f / d.z};
    if (aabb_intersect_ref(sv.root_min, sv.root_max, o, inv, RL_EPSILON, RL_F32_MAX)) h = flat_closest<true>(sv, sv.flat, sv.trav, o, d, quads);
    return h;
}

int emu_trace(const emu_scene *s, size_t n, const float *o, const float *d, uint32_t *prim, float *tuv) {
    for (size_t i = 0; i < n; i++) {
        HitRec h = trace_closest(s->sv, s->sv.nodes, s->sv.trav, V3{o[3 * i], o[3 * i + 1], o[3 * i + 2]}, V3{d[3 * i], d[3 * i + 1], d[3 * i + 2]});
        prim[i] = h.prim;
        if (tuv) {
            bool miss = h.prim == RL_MISS;
            tuv[3 * i] = miss ? 0.0f : h.t, tuv[3 * i + 1] = miss ? 0.0f : h.u, tuv[3 * i + 2] = miss ? 0.0f : h.v;
        }
    }
    return 0;
}
int emu_visible(const emu_scene *s, size_t n, const float *p0, const float *p1, uint8_t *out) {
    for (size_t i = 0; i < n; i++)
        out[i] = trace_visible(s->sv, s->sv.nodes, s->sv.trav, V3{p0[3 * i], p0[3 * i + 1], p0[3 * i + 2]}, V3{p1[3 * i], p1[3 * i + 1], p1[3 * i + 2]}) ? 1 : 0;
    return 0;
}
int emu_primary_hits(const emu_scene *s, uint32_t *prim, float *tuv) {
    uint32_t W = s->hs.img_w, H = s->hs.img_h;
    for (uint32_t y = 0; y < H; y++)
        for (uint32_t x = 0; x < W; x++) {
            V3 o, d;
            camera_generate(s->sv, (float)x + 0.5f, (float)y + 0.5f, &o, &d);
            size_t i = (size_t)y * W + x;
            HitRec h = trace_camera(s, x, y, o, d);
            prim[i] = h.prim;
            if (tuv) {
                bool miss = h.prim == RL_MISS;
                tuv[3 * i] = miss ? 0.0f : h.t, tuv[3 * i + 1] = miss ? 0.0f : h.u, tuv[3 * i + 2] = miss ? 0.0f : h.v;
            }
        }
    return 0;
}

// Device BSDF arithmetic on one material (rows built by the function build_host_scene uses).  A blend is passed as an array of
// three materials {blend, bsdf1, bsdf2} with blend_a = 1, blend_b = 2.
#define EMU_MAT_ROWS (3 * RL_MAT_F4)
static void emu_material_rows(const rl_material *mt, float4 rows[EMU_MAT_ROWS]) {
    material_rows(mt[0], rows, RL_MAT_F4, 2 * RL_MAT_F4);
    if (mt->kind == RL_BSDF_BLEND) {
        material_rows(mt[1], rows + RL_MAT_F4, 0, 0);
        material_rows(mt[2], rows + 2 * RL_MAT_F4, 0, 0);
    }
}
// returns 0 = None, 1 = SolidAngle pdf, 2 = Discrete pdf
int emu_bsdf_sample(const rl_material *mt, const float wi[3], float s0, float s1, float weight[3], float d[3], float *pdf) {
    float4 rows[EMU_MAT_ROWS];
    emu_material_rows(mt, rows);
    Material m = load_material(rows, 0);
    Col w;
    V3 wo;
    bool discrete;
    if (!bsdf_sample(m, V3{wi[0], wi[1], wi[2]}, s0, s1, &w, &wo, pdf, &discrete)) return 0;
    weight[0] = w.r, weight[1] = w.g, weight[2] = w.b;
    d[0] = wo.x, d[1] = wo.y, d[2] = wo.z;
    return discrete ? 2 : 1;
}
float emu_bsdf_pdf(const rl_material *mt, const float wi[3], const float wo[3]) {
    float4 rows[EMU_MAT_ROWS];
    emu_material_rows(mt, rows);
    return bsdf_pdf(load_material(rows, 0), V3{wi[0], wi[1], wi[2]}, V3{wo[0], wo[1], wo[2]});
}
void emu_bsdf_eval(const rl_material *mt, const float wi[3], const float wo[3], float out[3]) {
    float4 rows[EMU_MAT_ROWS];
    emu_material_rows(mt, rows);
    Col c = bsdf_eval(load_material(rows, 0), V3{wi[0], wi[1], wi[2]}, V3{wo[0], wo[1], wo[2]});
    out[0] = c.r, out[1] = c.g, out[2] = c.b;
}
// LightSamplerATS on the device side (rl_device.cuh: ats_*)
int emu_ats_sample(const emu_scene *s, float r, const float p[3], const float n[3], uint32_t *prim, float *pdf) {
    if (!s->sv.ats_nodes) return -1;
    *prim = ats_sample(s->sv, r, V3{p[0], p[1], p[2]}, V3{n[0], n[1], n[2]}, pdf);
    return 0;
}
int emu_ats_pdf(const emu_scene *s, uint32_t prim, const float p[3], const float n[3], int has_n, float *pdf) {
    if (!s->sv.ats_nodes || s->sv.ats_leaf_of_prim[prim] == 0xffffffffu) return -1;
    *pdf = ats_pdf(s->sv, prim, V3{p[0], p[1], p[2]}, V3{n[0], n[1], n[2]}, has_n != 0);
    return 0;
}
// EnvironmentLightColor::Texture on the device side (rl_device.cuh: env_*)
float emu_spec_atan2(float y, float x) { return spec_atan2f(y, x); }
float emu_spec_acos(float x) { return spec_acosf(x); }
int emu_env_eval_pdf(const emu_scene *s, const float d[3], float rgb[3], float *pdf) {
    if (!s->sv.env_w) return -1;
    Col c = env_eval(s->sv, V3{d[0], d[1], d[2]});
    rgb[0] = c.r, rgb[1] = c.g, rgb[2] = c.b;
    *pdf = env_pdf(s->sv, V3{d[0], d[1], d[2]});
    return 0;
}
int emu_env_sample(const emu_scene *s, float u0, float u1, float d[3], float rgb[3], float *pdf) {
    if (!s->sv.env_w) return -1;
    V3 dd;
    Col c;
    env_sample_direction(s->sv, u0, u1, &dd, &c, pdf);
    d[0] = dd.x, d[1] = dd.y, d[2] = dd.z;
    rgb[0] = c.r, rgb[1] = c.g, rgb[2] = c.b;
    return 0;
}
int emu_bsdf_flags(const rl_material *mt) {
    float4 rows[EMU_MAT_ROWS];
    emu_material_rows(mt, rows);
    Material m = load_material(rows, 0);
    return (mat_is_twosided(m) ? 1 : 0) | (mat_is_smooth(m) ? 2 : 0);
}

struct emu_stats {
    uint64_t samples, segments, shadow_rays, shadow_traced, shadow_visible, hits, max_depth_seen;
};

// The wavefront of rl_render, executed one path at a time (each path is independent).
int emu_render(const emu_scene *s, const rl_integrator_desc *I, uint32_t spp, uint64_t seed, uint32_t rank, uint32_t nranks, float *out_rgb,
               emu_stats *stats) {
    if (!s || !I || spp == 0) return RL_ERR_INVALID;
    const SceneView &sv = s->sv;
    const uint32_t W = s->hs.img_w, H = s->hs.img_h;
    IntegParams ip{};
    ip.kind = I->kind, ip.min_depth = I->min_depth, ip.max_depth = I->max_depth, ip.rr_depth = I->rr_depth;
    ip.strategy = I->strategy, ip.single_scattering = I->single_scattering;
    ip.nb_bsdf_samples = I->nb_bsdf_samples, ip.nb_light_samples = I->nb_light_samples;
    ip.ao_max_distance = I->ao_max_distance, ip.ao_normal_correction = I->ao_normal_correction;
    ip.seed_h = seed_hash(seed);
    ip.sample_base = 0, ip.npix = W * H, ip.img_w = W;
    emu_stats S{};
    std::memset(out_rgb, 0, sizeof(float) * 3 * (size_t)W * H);
    const float inv_spp = 1.0f / (float)spp;
    for (uint32_t py = 0; py < H; py++)
        for (uint32_t px = 0; px < W; px++) {
            if (nranks > 1 && ((px / 16 + py / 16) % nranks) != rank) continue;
            uint32_t pixel = py * W + px;
            float sum[3] = {0.0f, 0.0f, 0.0f};
            for (uint32_t sidx = 0; sidx < spp; sidx++) {
                // k_raygen
                Sampler smp = make_sampler(ip.seed_h, pixel, sidx, 0u);
                float jx = smp.next();
                float jy = smp.next();
                V3 o, d;
                camera_generate(sv, (float)px + jx, (float)py + jy, &o, &d);
                PathState st;
                st.T = Col{1.0f, 1.0f, 1.0f}, st.pdf_prev = 1.0f, st.path_id = 0, st.depth = 1, st.rng_n = smp.n;
                float L[3] = {0.0f, 0.0f, 0.0f};
                uint64_t iter = 0;
                if (I->kind == RL_INTEGRATOR_AO) {
                    // k_trace -> k_shade_direct1 (ao_begin / ao_sample) -> k_trace -> k_shade_direct2 (ao_finish)
                    S.segments++;
                    HitRec h = trace_camera(s, px, py, o, d);
                    if (h.prim != RL_MISS) S.hits++;
                    DirectCtx cx;
                    ao_begin(sv, ip, o, d, h, st.rng_n, pixel, sidx, &cx);
                    Col tot = Col{0.0f, 0.0f, 0.0f};
                    V3 dir;
                    if (cx.ok && ao_sample(ip, &cx, &dir)) {
                        S.segments++;
                        HitRec h2 = trace_closest(sv, sv.nodes, sv.trav, cx.its.p, dir);
                        if (h2.prim != RL_MISS) S.hits++;
                        Col c;
                        if (ao_finish(ip, h2, &c)) tot = tot + c;
                    }
                    sum[0] += tot.r, sum[1] += tot.g, sum[2] += tot.b;
                    S.max_depth_seen = std::max<uint64_t>(S.max_depth_seen, 2);
                    S.samples++;
                    continue;
                }
                if (I->kind == RL_INTEGRATOR_DIRECT) {
                    // k_trace -> k_shade_direct1 -> k_shadow -> k_trace -> k_shade_direct2; slots summed in order (k_accum)
                    std::vector<Col> slots(1 + ip.nb_light_samples + ip.nb_bsdf_samples, Col{0.0f, 0.0f, 0.0f});
                    S.segments++;
                    HitRec h = trace_camera(s, px, py, o, d);
                    if (h.prim != RL_MISS) S.hits++;
                    DirectCtx cx;
                    direct_begin(sv, ip, o, d, h, st.rng_n, pixel, sidx, &cx);
                    if (cx.ok || cx.env_primary) slots[0] = cx.emit;
                    for (uint32_t j = 0; j < ip.nb_light_samples && cx.ok; j++) {
                        V3 p1;
                        Col c;
                        bool valid;
                        bool sh = direct_light_sample(sv, &cx, &p1, &c, &valid);
                        if (valid) S.shadow_rays++;
                        if (sh) {
                            S.shadow_traced++;
                            if (trace_visible(sv, sv.nodes, sv.trav, cx.its.p, p1)) {
                                S.shadow_visible++;
                                slots[1 + j] = c;
                            }
                        }
                    }
                    for (uint32_t k = 0; k < ip.nb_bsdf_samples && cx.ok; k++) {
                        V3 dir;
                        Col w;
                        float pdf;
                        if (!direct_bsdf_sample(&cx, &dir, &w, &pdf)) continue;
                        S.segments++;
                        HitRec h2 = trace_closest(sv, sv.nodes, sv.trav, cx.its.p, dir);
                        if (h2.prim != RL_MISS) S.hits++;
                        Col c;
                        if (direct_finish(sv, ip, cx.its.p, dir, h2, w, pdf, &c, cx.its.n_s, sv.ats_nodes != nullptr)) slots[1 + ip.nb_light_samples + k] = c;
                    }
                    Col tot = slots[0];
                    for (size_t q = 1; q < slots.size(); q++) tot = tot + slots[q];
                    sum[0] += tot.r, sum[1] += tot.g, sum[2] += tot.b;
                    S.max_depth_seen = std::max<uint64_t>(S.max_depth_seen, 2);
                    S.samples++;
                    continue;
                }
                for (;;) {
                    iter++;
                    S.segments++;
                    HitRec h = iter == 1 ? trace_camera(s, px, py, o, d) : trace_closest(sv, sv.nodes, sv.trav, o, d); // k_trace
                    if (h.prim != RL_MISS) S.hits++;
                    StepOut so;
                    path_step(sv, ip, o, d, h, st, pixel, sidx, &so); // k_shade
                    if (so.has_add) L[0] += so.add.r, L[1] += so.add.g, L[2] += so.add.b;
                    if (so.nee_sampled) S.shadow_rays++;
                    if (so.shadow) { // k_shadow
                        S.shadow_traced++;
                        if (trace_visible(sv, sv.nodes, sv.trav, so.sh_p0, so.sh_p1)) {
                            S.shadow_visible++;
                            L[0] += so.sh_contrib.r, L[1] += so.sh_contrib.g, L[2] += so.sh_contrib.b;
                        }
                    }
                    if (!so.alive) break;
                    o = so.next_o, d = so.next_d, st = so.next;
                }
                S.max_depth_seen = std::max(S.max_depth_seen, iter);
                sum[0] += L[0], sum[1] += L[1], sum[2] += L[2]; // k_accum
                S.samples++;
            }
            out_rgb[3 * (size_t)pixel] = sum[0] * inv_spp, out_rgb[3 * (size_t)pixel + 1] = sum[1] * inv_spp, out_rgb[3 * (size_t)pixel + 2] = sum[2] * inv_spp; // k_finish
        }
    if (stats) *stats = S;
    return 0;
}

} // extern "C"
