// rl_scene_host.hpp -- flattens an rl_scene_desc into the float4 tables of rl_device.cuh.
//
// Host-side, once per scene (the reference does the same work in Mesh::new, geometry.rs:122-182,
// and Scene::build_emitters, scene.rs:53-123).  Sequential f32 accumulations (the area CDFs)
// are done here in the reference's order; compile with -ffp-contract=off.  The per-triangle
// traversal records and the LBVH are built on the device (rl_build.cuh / rl_kernels.cu).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "rl_b200.h"
#include "rl_ats_host.hpp"
#include "rl_device.cuh"

namespace rl {

struct HostScene {
    uint32_t ntris = 0, nmeshes = 0, n_emitters = 0;
    bool emit_var = false;         // a mesh light with EmissionType::HSV / Texture: emission depends on uv (kernels with KM bit 8)
    std::vector<float4> verts;     // 3 per prim
    std::vector<float4> shade;     // 4 per prim; [0].xyz (n_geo) filled by the device setup kernel
    std::vector<float4> mats;      // RL_MAT_F4 per mesh
    std::vector<float2> uvs;       // 3 per prim (zeros for meshes without uv); empty when no mesh has uv
    std::vector<float4> tex;       // 4 per texture: {color0, kind} {color1, line_width} {offset.xy, scale.xy} {width, height, texel offset, -}
    std::vector<float4> texels;    // bitmap pixels {r, g, b, -}
    std::vector<float4> emit_info; // 1 per emitter
    std::vector<float> emit_cdf;   // n_emitters + 1
    std::vector<float> area_cdf;   // concatenated
    float root_min[3], root_max[3]; // reference root box (padded per-triangle boxes, geometry.rs:423-439)
    float raw_min[3], raw_max[3];   // plain vertex bounds (Morton normalisation)
    float abs_max = 0;
    float s2c[16], c2w[16], cam_pos[3];
    // EnvironmentLight (constant): colour, bounding sphere (radius already x 1.1), emitter-selection pdf
    bool env_on = false;
    float env_color[3] = {0, 0, 0}, bs_center[3] = {0, 0, 0}, bs_radius = 0.0f, env_pdf_sel = 0.0f;
    // EnvironmentLightColor::Texture: the image (in texels[] at env_texel_off) and its Distribution2D (rl_device.cuh: SceneView::env_dist)
    uint32_t env_w = 0, env_h = 0, env_texel_off = 0;
    std::vector<float> env_dist;
    float env_func_int = 0.0f;
    // LightSamplerATS (rl_ats_host.hpp): filled by build_host_scene when desc->use_ats
    std::vector<float4> ats_nodes;
    std::vector<uint32_t> ats_leaf_of_prim;
    uint32_t ats_root = 0, ats_depth = 0;
    uint32_t img_w = 0, img_h = 0;
};

inline float4 f4(float x, float y, float z, float w) {
    float4 r;
    r.x = x, r.y = y, r.z = z, r.w = w;
    return r;
}

// Distribution1DConstruct::normalize, math.rs:418-441
inline float dist1d_normalize(const std::vector<float> &elements, std::vector<float> &cdf) {
    cdf.clear();
    float cur = 0.0f;
    for (float e : elements) {
        cdf.push_back(cur);
        cur += e / (float)elements.size();
    }
    cdf.push_back(cur);
    if (cur != 0.0f)
        for (float &x : cdf) x /= cur;
    cdf.back() = 1.0f;
    return cur; // func_int
}

// BSDF parameters the device supports (sub = one of the two parts of a BSDFBlend: rough, constant colours, not a blend itself)
inline bool material_ok(const rl_material &mt, const rl_scene_desc *desc, bool sub) {
    if (mt.kind == RL_BSDF_BLEND) {
        if (sub || !desc->submaterials || mt.blend_a == 0 || mt.blend_b == 0 || mt.blend_a > desc->nsubmaterials || mt.blend_b > desc->nsubmaterials) return false;
        if (!(mt.blend_weight >= 0.0f && mt.blend_weight <= 1.0f)) return false;
        return material_ok(desc->submaterials[mt.blend_a - 1], desc, true) && material_ok(desc->submaterials[mt.blend_b - 1], desc, true);
    }
    const bool has_mf = mt.kind == RL_BSDF_METAL || mt.kind == RL_BSDF_SUBSTRATE;
    if (mt.kind > RL_BSDF_SUBSTRATE || (has_mf && mt.microfacet > RL_MICROFACET_BECKMANN) || (has_mf && mt.microfacet != RL_MICROFACET_NONE && !(mt.alpha > 0.0f)) ||
        (mt.kind == RL_BSDF_GLASS && mt.ior == 0.0f))
        return false;
    if (mt.kd_texture > desc->ntextures || mt.ks_texture > desc->ntextures || mt.kt_texture > desc->ntextures || mt.eta_texture > desc->ntextures ||
        mt.k_texture > desc->ntextures || (mt.kd_texture != 0 && mt.kind != RL_BSDF_DIFFUSE && mt.kind != RL_BSDF_PHONG && mt.kind != RL_BSDF_SUBSTRATE) ||
        (mt.ks_texture != 0 && mt.kind == RL_BSDF_DIFFUSE) || (mt.kt_texture != 0 && mt.kind != RL_BSDF_GLASS) ||
        ((mt.eta_texture != 0 || mt.k_texture != 0) && mt.kind != RL_BSDF_METAL))
        return false;
    if (sub) { // blend.rs:17 asserts !is_smooth() on both parts
        const bool smooth = mt.kind == RL_BSDF_GLASS || (has_mf && mt.microfacet == RL_MICROFACET_NONE);
        if (smooth || mt.kd_texture || mt.ks_texture || mt.kt_texture || mt.eta_texture || mt.k_texture) return false;
    }
    return true;
}
// The BSDF part of the RL_MAT_F4 rows of one material (rl_device.cuh: load_material); the caller adds the emitter fields (row 2, row 3 .y .z).
// delta_a / delta_b: distance in float4 rows from this material's first row to the first row of the two parts of a blend.
inline void material_rows(const rl_material &mt, float4 rows[6], int delta_a, int delta_b) {
    const float *ca = mt.kind == RL_BSDF_METAL ? mt.eta : (mt.kind == RL_BSDF_GLASS ? mt.kt : mt.kd);
    const bool has_mf = mt.kind == RL_BSDF_METAL || mt.kind == RL_BSDF_SUBSTRATE;
    const bool blend = mt.kind == RL_BSDF_BLEND;
    rows[0] = f4(ca[0], ca[1], ca[2], u2f(mt.kind));
    rows[1] = f4(mt.ks[0], mt.ks[1], mt.ks[2], has_mf ? mt.alpha : mt.exponent);
    rows[2] = f4(0.0f, 0.0f, 0.0f, u2f(0u));
    rows[3] = f4(mt.kind == RL_BSDF_GLASS ? mt.ior : (blend ? mt.blend_weight : mt.weight_specular), 0.0f, 0.0f, u2f(has_mf ? mt.microfacet : 0u));
    // row 4: {metal k, glass 1/eta (BSDFGlass::eta(): inv_eta = 1.0 / eta)} | blend: {row delta of bsdf1, row delta of bsdf2} as int bits
    rows[4] = blend ? f4(u2f((uint32_t)delta_a), u2f((uint32_t)delta_b), 0.0f, 0.0f) : f4(mt.k[0], mt.k[1], mt.k[2], mt.kind == RL_BSDF_GLASS ? 1.0f / mt.ior : 0.0f);
    // row 5: textures of the three colour slots as uint bits: {slot a = kd | metal eta | glass kt, slot b = ks, slot c = metal k, -}
    const uint32_t ta = mt.kind == RL_BSDF_METAL ? mt.eta_texture : (mt.kind == RL_BSDF_GLASS ? mt.kt_texture : mt.kd_texture);
    rows[5] = blend ? f4(0.0f, 0.0f, 0.0f, 0.0f) : f4(u2f(ta), u2f(mt.ks_texture), u2f(mt.kind == RL_BSDF_METAL ? mt.k_texture : 0u), 0.0f);
}
// Mesh::emit(uv) for EmissionType::HSV / Texture on the host (the light tree's proxies, emitter.rs:742-756); the operations of
// rl_device.cuh: mesh_emit.  `out` = three floats.
inline void host_mesh_emit(const rl_scene_desc *desc, const rl_mesh_desc &m, float u, float v, float *out) {
    const float scale = m.emission[0];
    Col c;
    if (m.emission_kind == RL_EMISSION_HSV) {
        const float x = fmodf(fabsf(u), 1.0f); // uv.x.abs() % 1.0
        c = Col{x * 1.0f + (1.0f - x) * 0.0f, x * 0.0f + (1.0f - x) * 1.0f, x * 0.0f + (1.0f - x) * 0.0f};
    } else { // img.pixel_uv(uv), structure.rs:434-453
        const rl_texture &t = desc->textures[m.emission_texture - 1];
        auto modulo1 = [](float a) { return fmodf(fmodf(a, 1.0f) + 1.0f, 1.0f); };
        auto as_usize = [](float a) -> uint64_t { return !(a > 0.0f) ? 0ull : (a >= 18446744073709551616.0f ? ~0ull : (uint64_t)a); };
        const uint64_t x = as_usize(modulo1(u) * (float)t.width), y = as_usize(modulo1(v) * (float)t.height), i = (uint64_t)t.width * y + x;
        c = i >= (uint64_t)t.width * t.height ? Col{0.0f, 0.0f, 0.0f} : Col{t.pixels[3 * i], t.pixels[3 * i + 1], t.pixels[3 * i + 2]};
    }
    c = mul_checked(c, scale);
    out[0] = c.r, out[1] = c.g, out[2] = c.b;
}
inline bool build_host_scene(const rl_scene_desc *desc, HostScene &hs, std::string &err) {
    if (!desc || !desc->meshes || desc->nmeshes == 0) {
        err = "empty scene";
        return false;
    }
    if (desc->has_volume) {
        err = "scene.volume must be None on this path";
        return false;
    }
    if (desc->has_environment > 2 || (desc->has_environment == 2 && (desc->environment_texture == 0 || desc->environment_texture > desc->ntextures ||
                                                                         !desc->textures || desc->textures[desc->environment_texture - 1].kind != RL_TEX_BITMAP))) {
        err = "has_environment == 2 needs environment_texture = 1 + index of a bitmap texture";
        return false;
    }
    if (desc->camera.width == 0 || desc->camera.height == 0) {
        err = "empty image";
        return false;
    }
    hs = HostScene();
    hs.nmeshes = desc->nmeshes;
    hs.img_w = desc->camera.width;
    hs.img_h = desc->camera.height;
    std::memcpy(hs.s2c, desc->camera.sample_to_camera, 64);
    std::memcpy(hs.c2w, desc->camera.to_world, 64);
    { // Camera::position = to_world.transform_point(0,0,0), camera.rs:140-142
        float h[4];
        m4_mul_v4(hs.c2w, 0.0f, 0.0f, 0.0f, 1.0f, h);
        float iw = 1.0f / h[3];
        hs.cam_pos[0] = h[0] * iw, hs.cam_pos[1] = h[1] * iw, hs.cam_pos[2] = h[2] * iw;
    }
    for (int a = 0; a < 3; a++) {
        hs.root_min[a] = hs.raw_min[a] = RL_F32_MAX;
        hs.root_max[a] = hs.raw_max[a] = -RL_F32_MAX;
    }
    struct EmitterTmp {
        uint32_t mesh, first_prim, ntris, cdf_off; // mesh lights
        float flux_max;
        uint32_t light_kind = 0xffffffffu; // rl_light_kind for non-mesh emitters
        float intensity[3] = {0, 0, 0}, v[3] = {0, 0, 0}, radius = 0.0f;
    };
    std::vector<EmitterTmp> emitters;
    bool any_uv = false;
    for (uint32_t mi = 0; mi < desc->nmeshes; mi++) any_uv = any_uv || desc->meshes[mi].UV != nullptr;
    // textures (BSDFColor::{Bitmap, Checkerbord, Grid}) -> 4 rows each + the texel array
    if (desc->ntextures > 0 && !desc->textures) {
        err = "ntextures > 0 but textures is null";
        return false;
    }
    for (uint32_t ti = 0; ti < desc->ntextures; ti++) {
        const rl_texture &t = desc->textures[ti];
        if (t.kind < RL_TEX_BITMAP || t.kind > RL_TEX_GRID || (t.kind == RL_TEX_BITMAP && (!t.pixels || t.width == 0 || t.height == 0))) {
            err = "bad texture";
            return false;
        }
        uint32_t off = (uint32_t)hs.texels.size();
        hs.tex.push_back(f4(t.color0[0], t.color0[1], t.color0[2], u2f(t.kind)));
        hs.tex.push_back(f4(t.color1[0], t.color1[1], t.color1[2], t.line_width));
        hs.tex.push_back(f4(t.offset[0], t.offset[1], t.scale[0], t.scale[1]));
        hs.tex.push_back(f4(u2f(t.width), u2f(t.height), u2f(off), 0.0f));
        if (t.kind == RL_TEX_BITMAP)
            for (size_t i = 0; i < (size_t)t.width * t.height; i++) hs.texels.push_back(f4(t.pixels[3 * i], t.pixels[3 * i + 1], t.pixels[3 * i + 2], 0.0f));
    }
    std::vector<float> mesh_inv_area(desc->nmeshes, 0.0f);
    uint32_t first = 0;
    for (uint32_t mi = 0; mi < desc->nmeshes; mi++) {
        const rl_mesh_desc &m = desc->meshes[mi];
        if (!m.P || !m.idx || m.ntris == 0 || m.nverts == 0) {
            err = "mesh without geometry";
            return false;
        }
        if (!material_ok(m.mat, desc, false)) {
            err = "unsupported BSDF kind or parameters";
            return false;
        }
        if (m.emission_kind > RL_EMISSION_TEXTURE) {
            err = "unknown emission kind";
            return false;
        }
        if (m.emission_kind >= RL_EMISSION_HSV) { // Mesh::emit unwraps the uv of the hit / of the sampled point (geometry.rs:197, 204)
            if (!m.UV) {
                err = "HSV / textured emission on a mesh without uv coordinates (the reference panics: uv.unwrap(), geometry.rs:197)";
                return false;
            }
            if (m.emission_kind == RL_EMISSION_TEXTURE &&
                (m.emission_texture == 0 || m.emission_texture > desc->ntextures || desc->textures[m.emission_texture - 1].kind != RL_TEX_BITMAP)) {
                err = "emission_texture must be 1 + index of a bitmap texture";
                return false;
            }
            hs.emit_var = true;
        }
        std::vector<float> areas;
        for (uint32_t t = 0; t < m.ntris; t++) {
            uint32_t id[3] = {m.idx[3 * t], m.idx[3 * t + 1], m.idx[3 * t + 2]};
            V3 v[3];
            for (int k = 0; k < 3; k++) {
                if (id[k] >= m.nverts) {
                    err = "triangle index out of range";
                    return false;
                }
                v[k] = V3{m.P[3 * id[k]], m.P[3 * id[k] + 1], m.P[3 * id[k] + 2]};
                hs.verts.push_back(f4(v[k].x, v[k].y, v[k].z, 0.0f));
            }
            // shading record: n_geo slot (device fills xyz), mesh id, vertex normals; flags: bit 0 normals, bit 1 uv
            const uint32_t flags = (m.N ? 1u : 0u) | (m.UV ? 2u : 0u);
            hs.shade.push_back(f4(0.0f, 0.0f, 0.0f, u2f(mi)));
            for (int k = 0; k < 3; k++) {
                if (m.N) hs.shade.push_back(f4(m.N[3 * id[k]], m.N[3 * id[k] + 1], m.N[3 * id[k] + 2], k == 0 ? u2f(flags) : 0.0f));
                else hs.shade.push_back(f4(0.0f, 0.0f, 0.0f, k == 0 ? u2f(flags) : 0.0f));
            }
            if (any_uv)
                for (int k = 0; k < 3; k++) {
                    float2 q;
                    q.x = m.UV ? m.UV[2 * id[k]] : 0.0f, q.y = m.UV ? m.UV[2 * id[k] + 1] : 0.0f;
                    hs.uvs.push_back(q);
                }
            // bounds: compute_aabb_tri (geometry.rs:423-439) for the reference root box
            float lo[3], hi[3];
            for (int a = 0; a < 3; a++) {
                float c0 = a == 0 ? v[0].x : (a == 1 ? v[0].y : v[0].z);
                float c1 = a == 0 ? v[1].x : (a == 1 ? v[1].y : v[1].z);
                float c2 = a == 0 ? v[2].x : (a == 1 ? v[2].y : v[2].z);
                lo[a] = fminf(fminf(c0, c1), c2);
                hi[a] = fmaxf(fmaxf(c0, c1), c2);
                hs.raw_min[a] = fminf(hs.raw_min[a], lo[a]);
                hs.raw_max[a] = fmaxf(hs.raw_max[a], hi[a]);
                if (hi[a] - lo[a] < RL_EPSILON) {
                    hi[a] += RL_EPSILON;
                    lo[a] -= RL_EPSILON;
                }
                hs.root_min[a] = fminf(hs.root_min[a], lo[a]);
                hs.root_max[a] = fmaxf(hs.root_max[a], hi[a]);
            }
            areas.push_back(magnitude(cross(v[1] - v[0], v[2] - v[0])) * 0.5f); // Mesh::new, geometry.rs:136
        }
        // Mesh.cdf, total(), pdf() = 1/total  (geometry.rs:179,223-225; math.rs:484-486)
        std::vector<float> cdf;
        float func_int = dist1d_normalize(areas, cdf);
        float total = func_int * (float)(cdf.size() - 1);
        mesh_inv_area[mi] = 1.0f / total;
        if (m.emission_kind != 0) {
            EmitterTmp e;
            e.mesh = mi, e.first_prim = first, e.ntris = m.ntris, e.cdf_off = (uint32_t)hs.area_cdf.size();
            // Mesh::flux = total * Le * PI, emitter.rs:591-599; channel_max for the emitter CDF, scene.rs:103-111
            // (HSV / Texture: e = Color::value(scale), "TODO" in the reference)
            const bool var = m.emission_kind >= RL_EMISSION_HSV;
            float fr = (m.emission[0] * total), fg = ((var ? m.emission[0] : m.emission[1]) * total), fb = ((var ? m.emission[0] : m.emission[2]) * total);
            Col fl = mul_checked(Col{fr, fg, fb}, RL_PI);
            e.flux_max = channel_max(fl);
            emitters.push_back(e);
            hs.area_cdf.insert(hs.area_cdf.end(), cdf.begin(), cdf.end());
        }
        first += m.ntris;
    }
    hs.ntris = first;
    // Non-mesh emitters follow the mesh lights (scene.rs:85-96).  Scene.bsphere (scene.rs:54-60): union of
    // Mesh::compute_aabb over ALL vertices of every mesh (geometry.rs:441-456) and the camera position, to_sphere
    // (structure.rs:871-877); DirectionalLight::preprocess enlarges the radius by 1.1 (emitter.rs:106-109).
    if (desc->nlights > 0 && !desc->lights) {
        err = "nlights > 0 but lights is null";
        return false;
    }
    if (desc->nlights > 0 || desc->has_environment) {
        float bmin[3] = {RL_F32_MAX, RL_F32_MAX, RL_F32_MAX}, bmax[3] = {-RL_F32_MAX, -RL_F32_MAX, -RL_F32_MAX};
        for (uint32_t mi = 0; mi < desc->nmeshes; mi++) {
            const rl_mesh_desc &m = desc->meshes[mi];
            float lo[3] = {RL_F32_MAX, RL_F32_MAX, RL_F32_MAX}, hi[3] = {-RL_F32_MAX, -RL_F32_MAX, -RL_F32_MAX};
            for (uint32_t k = 0; k < m.nverts; k++)
                for (int a = 0; a < 3; a++) lo[a] = fminf(lo[a], m.P[3 * k + a]), hi[a] = fmaxf(hi[a], m.P[3 * k + a]);
            for (int a = 0; a < 3; a++) {
                if (hi[a] - lo[a] < RL_EPSILON) hi[a] += RL_EPSILON, lo[a] -= RL_EPSILON;
                bmin[a] = fminf(bmin[a], lo[a]), bmax[a] = fmaxf(bmax[a], hi[a]);
            }
        }
        for (int a = 0; a < 3; a++) bmin[a] = fminf(bmin[a], hs.cam_pos[a]), bmax[a] = fmaxf(bmax[a], hs.cam_pos[a]);
        V3 size = V3{bmax[0] - bmin[0], bmax[1] - bmin[1], bmax[2] - bmin[2]};
        V3 c = size * 0.5f + V3{bmin[0], bmin[1], bmin[2]}; // AABB::center, structure.rs:844-846
        float radius = magnitude(c - V3{bmax[0], bmax[1], bmax[2]});
        if (desc->has_environment) { // scene.rs:69-81: the environment follows the mesh lights; flux = PI r^2 c (emitter.rs:512-516)
            EmitterTmp e{};
            e.light_kind = 2u;
            e.radius = radius * 1.1f; // EnvironmentLight::preprocess, emitter.rs:436-439
            for (int a = 0; a < 3; a++) e.intensity[a] = desc->has_environment == 2 ? 0.0f : desc->environment[a];
            e.v[0] = c.x, e.v[1] = c.y, e.v[2] = c.z;
            e.flux_max = channel_max(mul_plain(RL_PI * (e.radius * e.radius), Col{e.intensity[0], e.intensity[1], e.intensity[2]}));
            if (desc->has_environment == 2) { // EnvironmentLightColor::new_texture (emitter.rs:341-353) + Distribution2D::from_bitmap (math.rs:495-521)
                const uint32_t ti = desc->environment_texture - 1;
                const uint32_t W = f2u(hs.tex[4 * ti + 3].x), H = f2u(hs.tex[4 * ti + 3].y), off = f2u(hs.tex[4 * ti + 3].z);
                hs.env_w = W, hs.env_h = H, hs.env_texel_off = off;
                std::vector<float> marginal_elems, cond_cdfs, cond_funcs, row(W), cdf;
                for (uint32_t y = 0; y < H; y++) {
                    float sw, cw; // w = ((y + 0.5) * PI / size.y).sin(): the spec sine, like every sine of the device
                    spec_sincos(((float)y + 0.5f) * RL_PI / (float)H, &sw, &cw);
                    for (uint32_t x = 0; x < W; x++) {
                        const float4 px = hs.texels[off + (size_t)y * W + x];
                        const float r = px.x * sw, g = px.y * sw, b = px.z * sw;            // image_pdf pixel *= w (MulAssign<f32>, structure.rs:225-231)
                        row[x] = r * 0.212671f + g * 0.715160f + b * 0.072169f;              // Color::luminance, structure.rs:173-176
                    }
                    marginal_elems.push_back(dist1d_normalize(row, cdf));
                    cond_cdfs.insert(cond_cdfs.end(), cdf.begin(), cdf.end());
                    cond_funcs.insert(cond_funcs.end(), row.begin(), row.end());
                }
                hs.env_func_int = dist1d_normalize(marginal_elems, cdf);
                hs.env_dist = cdf;
                hs.env_dist.insert(hs.env_dist.end(), cond_cdfs.begin(), cond_cdfs.end());
                hs.env_dist.insert(hs.env_dist.end(), cond_funcs.begin(), cond_funcs.end());
                e.flux_max = RL_PI * (e.radius * e.radius) * hs.env_func_int; // Color::value(PI r^2 func_int), emitter.rs:517-523
            }
            emitters.push_back(e);
            hs.env_on = true;
            for (int a = 0; a < 3; a++) hs.env_color[a] = e.intensity[a];
            hs.bs_center[0] = c.x, hs.bs_center[1] = c.y, hs.bs_center[2] = c.z;
            hs.bs_radius = e.radius;
        }
        for (uint32_t li = 0; li < desc->nlights; li++) {
            const rl_light_desc &l = desc->lights[li];
            EmitterTmp e{};
            e.light_kind = l.kind;
            for (int a = 0; a < 3; a++) e.intensity[a] = l.intensity[a], e.v[a] = l.v[a];
            Col I = Col{l.intensity[0], l.intensity[1], l.intensity[2]};
            if (l.kind == RL_LIGHT_POINT) {
                e.flux_max = channel_max(mul_checked(mul_checked(I, 4.0f), RL_PI)); // emitter.rs:239-241
            } else if (l.kind == RL_LIGHT_DIRECTIONAL) { // kinds 0 / 1 = rl_light_kind, 2 = the environment (above)
                e.radius = radius * 1.1f;
                float area = RL_PI * (e.radius * e.radius);
                e.flux_max = channel_max(mul_plain(area, I)); // emitter.rs:164-168
            } else {
                err = "unknown light kind";
                return false;
            }
            emitters.push_back(e);
        }
    }
    hs.n_emitters = (uint32_t)emitters.size();
    std::vector<float> pdf_sel(desc->nmeshes, 0.0f);
    if (!emitters.empty()) {
        std::vector<float> fl;
        for (auto &e : emitters) fl.push_back(e.flux_max);
        dist1d_normalize(fl, hs.emit_cdf);
        // two float4 per emitter.  Mesh light: {mesh, first_prim, ntris, cdf_offset} {-}.
        // Point / directional: {0xfffffff0 | rl_light_kind, intensity.rgb} {position | direction, bounding-sphere radius}
        for (size_t i = 0; i < emitters.size(); i++) {
            const EmitterTmp &e = emitters[i];
            if (e.light_kind == 0xffffffffu) {
                pdf_sel[e.mesh] = hs.emit_cdf[i + 1] - hs.emit_cdf[i];
                hs.emit_info.push_back(f4(u2f(e.mesh), u2f(e.first_prim), u2f(e.ntris), u2f(e.cdf_off)));
                hs.emit_info.push_back(f4(0, 0, 0, 0));
            } else {
                hs.emit_info.push_back(f4(u2f(0xfffffff0u | e.light_kind), e.intensity[0], e.intensity[1], e.intensity[2]));
                hs.emit_info.push_back(f4(e.v[0], e.v[1], e.v[2], e.radius));
                if (e.light_kind == 2u) hs.env_pdf_sel = hs.emit_cdf[i + 1] - hs.emit_cdf[i];
            }
        }
    } else {
        hs.emit_cdf = {0.0f, 1.0f};
        hs.emit_info.push_back(f4(0, 0, 0, 0));
        hs.emit_info.push_back(f4(0, 0, 0, 0));
        hs.area_cdf = {0.0f, 1.0f};
    }
    for (uint32_t mi = 0; mi < desc->nmeshes; mi++) {
        const rl_mesh_desc &m = desc->meshes[mi];
        // rows of rl_device.cuh: load_material; the submaterials of blends follow the meshes
        float4 rows[6];
        const int base = 6 * ((int)desc->nmeshes - (int)mi);
        material_rows(m.mat, rows, base + 6 * ((int)m.mat.blend_a - 1), base + 6 * ((int)m.mat.blend_b - 1));
        // {Le.rgb, 1} | HSV {scale, -, -, 2} | Texture {scale, texture index (bits), -, 3}: rl_device.cuh: mesh_emit
        if (m.emission_kind >= RL_EMISSION_HSV) rows[2] = f4(m.emission[0], u2f(m.emission_kind == RL_EMISSION_TEXTURE ? m.emission_texture - 1u : 0u), 0.0f, u2f(m.emission_kind));
        else rows[2] = f4(m.emission_kind ? m.emission[0] : 0.0f, m.emission_kind ? m.emission[1] : 0.0f, m.emission_kind ? m.emission[2] : 0.0f, u2f(m.emission_kind ? 1u : 0u));
        rows[3].y = mesh_inv_area[mi], rows[3].z = pdf_sel[mi];
        hs.mats.insert(hs.mats.end(), rows, rows + 6);
    }
    for (uint32_t si = 0; si < desc->nsubmaterials; si++) {
        float4 rows[6];
        material_rows(desc->submaterials[si], rows, 0, 0);
        hs.mats.insert(hs.mats.end(), rows, rows + 6);
    }
    float am = 0.0f;
    for (int a = 0; a < 3; a++) am = fmaxf(am, fmaxf(fabsf(hs.raw_min[a]), fabsf(hs.raw_max[a])));
    hs.abs_max = am;
    if (desc->use_ats) { // Scene::build_emitters(true) -> LightSamplerATS::new (emitter.rs:1290-1317): every emitter must be a surface
        if (desc->nlights > 0 || desc->has_environment) {
            err = "use_ats: the light tree takes mesh emitters only (assert!(e.is_surface()), emitter.rs:1292-1294)";
            return false;
        }
        std::vector<uint8_t> emissive(hs.ntris, 0);
        std::vector<float> le(3 * (size_t)hs.ntris, 0.0f);
        uint32_t p = 0;
        for (uint32_t mi = 0; mi < desc->nmeshes; mi++)
            for (uint32_t t = 0; t < desc->meshes[mi].ntris; t++, p++)
                if (desc->meshes[mi].emission_kind) {
                    emissive[p] = 1;
                    const rl_mesh_desc &m = desc->meshes[mi];
                    if (m.emission_kind >= RL_EMISSION_HSV) { // self.emit(&uv) at the centroid uv, (uv0 + uv1 + uv2) / 3.0 (emitter.rs:742-756)
                        const uint32_t *ix = m.idx + 3 * (size_t)t;
                        const float cu = ((m.UV[2 * ix[0]] + m.UV[2 * ix[1]]) + m.UV[2 * ix[2]]) / 3.0f;
                        const float cv = ((m.UV[2 * ix[0] + 1] + m.UV[2 * ix[1] + 1]) + m.UV[2 * ix[2] + 1]) / 3.0f;
                        host_mesh_emit(desc, m, cu, cv, &le[3 * (size_t)p]);
                    } else
                        for (int a = 0; a < 3; a++) le[3 * (size_t)p + a] = m.emission[a];
                }
        AtsTree tree;
        if (!build_ats_tree(hs.verts, hs.ntris, emissive, le, tree, err)) return false;
        hs.ats_nodes = tree.nodes, hs.ats_leaf_of_prim = tree.leaf_of_prim, hs.ats_root = tree.root, hs.ats_depth = tree.depth;
    }
    return true;
}

} // namespace rl
