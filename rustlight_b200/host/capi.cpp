// capi.cpp -- C entry points of the host layer (librl_host.so) so that tests, bench.py and the
// Python mirror can load scenes and build cameras through the same C++ code the CLI uses.
// No rendering happens here; see include/rl_b200.h for the device library.
#include <cstring>

#include "rl_host.hpp"

using namespace rlh;

extern "C" {

struct rlh_scene {
    Scene scene;
};

static void set_err(char *err, size_t errlen, const std::string &msg) {
    if (err && errlen) {
        std::strncpy(err, msg.c_str(), errlen - 1);
        err[errlen - 1] = 0;
    }
}

// SceneLoaderManager::default().load(filename, use_shading_normal), src/scene_loader.rs:28-45
rlh_scene *rlh_load_scene(const char *filename, int use_shading_normal, char *err, size_t errlen) {
    try {
        auto *s = new rlh_scene;
        s->scene = SceneLoaderManager().load(filename, use_shading_normal != 0);
        return s;
    } catch (const std::exception &e) {
        set_err(err, errlen, e.what());
        return nullptr;
    }
}

// fmt = "pbrt" | "json"
rlh_scene *rlh_load_scene_string(const char *text, const char *fmt, int use_shading_normal, char *err, size_t errlen) {
    try {
        auto *s = new rlh_scene;
        std::string f(fmt ? fmt : "");
        if (f == "pbrt") s->scene = PBRTSceneLoader().load_string(text, use_shading_normal != 0);
        else if (f == "json") s->scene = JSONSceneLoader().load_string(text, use_shading_normal != 0);
        else if (f == "xml") s->scene = MTSSceneLoader().load_string(text, use_shading_normal != 0);
        else {
            delete s;
            throw Error("Impossible to found scene loader for " + f + " extension");
        }
        return s;
    } catch (const std::exception &e) {
        set_err(err, errlen, e.what());
        return nullptr;
    }
}

void rlh_scene_free(rlh_scene *s) { delete s; }

const rl_scene_desc *rlh_scene_desc(rlh_scene *s) { return s ? s->scene.desc() : nullptr; }

uint32_t rlh_scene_nb_meshes(const rlh_scene *s) { return (uint32_t)s->scene.meshes.size(); }
uint64_t rlh_scene_nb_triangles(const rlh_scene *s) { return (uint64_t)s->scene.nb_triangles(); }

// Camera::scale_image, src/camera.rs:73-78 (CLI `-s`)
void rlh_scene_scale_image(rlh_scene *s, float scale) { s->scene.camera.scale_image(scale); }

// Replace the film resolution and rebuild the camera (a scene-file edit, e.g. config C5's
// 1920x1080 Film; NOT the `-s` flag).
int rlh_scene_set_resolution(rlh_scene *s, uint32_t w, uint32_t h, char *err, size_t errlen) {
    try {
        const Camera &c = s->scene.camera;
        s->scene.camera = Camera::create(w, h, c.fov_axis, c.fov_deg, c.to_world, c.flip);
        return 0;
    } catch (const std::exception &e) {
        set_err(err, errlen, e.what());
        return -1;
    }
}

int rlh_scene_set_material(rlh_scene *s, uint32_t mesh, const rl_material *m) {
    if (!s || !m || mesh >= s->scene.meshes.size() || m->kind == RL_BSDF_BLEND) return -1; // blends: rlh_scene_set_material_blend
    s->scene.meshes[mesh]->bsdf.m = *m;
    s->scene.meshes[mesh]->bsdf.subs.clear();
    return 0;
}
// BSDFBlend { bsdf1: a, bsdf2: b, weight } (bsdfs/blend.rs)
int rlh_scene_set_material_blend(rlh_scene *s, uint32_t mesh, const rl_material *a, const rl_material *b, float weight) {
    if (!s || !a || !b || mesh >= s->scene.meshes.size()) return -1;
    try {
        Material ma, mb;
        ma.m = *a, mb.m = *b;
        s->scene.meshes[mesh]->bsdf = Material::blend(ma, mb, weight);
        return 0;
    } catch (const std::exception &) {
        return -1;
    }
}

// Material::phong incl. weight_specular, src/bsdfs/mod.rs:518-523
int rlh_material_phong(const float kd[3], const float ks[3], float exponent, rl_material *out) {
    try {
        *out = Material::phong(Color{kd[0], kd[1], kd[2]}, Color{ks[0], ks[1], ks[2]}, exponent).m;
        return 0;
    } catch (const std::exception &) {
        return -1;
    }
}

// Scene.emitters: kind is an rl_light_kind; v = position (point) or direction (directional, normalised here)
int rlh_scene_add_light(rlh_scene *s, uint32_t kind, const float intensity[3], const float v[3]) {
    try {
        Color I{intensity[0], intensity[1], intensity[2]};
        if (kind == RL_LIGHT_POINT) s->scene.add_point_light(I, v[0], v[1], v[2]);
        else if (kind == RL_LIGHT_DIRECTIONAL) s->scene.add_directional_light(I, v[0], v[1], v[2]);
        else return -1;
        return 0;
    } catch (const std::exception &) {
        return -1;
    }
}

// Scene.emitter_environment = EnvironmentLight with a constant colour
void rlh_scene_set_environment(rlh_scene *s, const float rgb[3]) { s->scene.set_environment(Color{rgb[0], rgb[1], rgb[2]}); }
// Scene::build_emitters(build_ats): light sampling through LightSamplerATS (`-x ats`)
void rlh_scene_set_ats(rlh_scene *s, int on) { s->scene.use_ats = on != 0; }
// `-x hvs-light` (tex_id == 0) / `-x texture-light` (tex_id = a bitmap texture id): examples/cli.rs:410-429
int rlh_scene_override_lights(rlh_scene *s, uint32_t tex_id) {
    try {
        if (tex_id == 0) s->scene.override_lights_hsv();
        else s->scene.override_lights_texture(tex_id);
        return 0;
    } catch (const std::exception &) {
        return -1;
    }
}
// EnvironmentLightColor::new_texture(image): `tex_id` = a bitmap texture id from rlh_scene_add_texture / rlh_scene_add_texture_file
int rlh_scene_set_environment_texture(rlh_scene *s, uint32_t tex_id) {
    try {
        s->scene.set_environment_texture(tex_id);
        return 0;
    } catch (const std::exception &) {
        return -1;
    }
}

// BSDFColor::{Bitmap, Checkerbord, Grid}: returns the 1-based id for rl_material.kd_texture, 0 on error.
// kind: rl_texture_kind; bitmap: (w, h, rgb[3*w*h]); checkerboard / grid: params = {color0[3], color1[3], offset[2], scale[2], line_width}
uint32_t rlh_scene_add_texture(rlh_scene *s, uint32_t kind, uint32_t w, uint32_t h, const float *rgb, const float *params) {
    try {
        if (kind == RL_TEX_BITMAP) return s->scene.add_texture(Texture::bitmap(w, h, std::vector<float>(rgb, rgb + (size_t)3 * w * h)));
        Color c0{params[0], params[1], params[2]}, c1{params[3], params[4], params[5]};
        if (kind == RL_TEX_CHECKERBOARD) return s->scene.add_texture(Texture::checkerboard(c0, c1, params[6], params[7], params[8], params[9]));
        if (kind == RL_TEX_GRID) return s->scene.add_texture(Texture::grid(c0, c1, params[10], params[6], params[7], params[8], params[9]));
        return 0;
    } catch (const std::exception &) {
        return 0;
    }
}
uint32_t rlh_scene_add_texture_file(rlh_scene *s, const char *filename) {
    try {
        return s->scene.add_texture(Texture::bitmap_file(filename));
    } catch (const std::exception &) {
        return 0;
    }
}

// Material::metal / glass / substrate (bsdfs/metal.rs, glass.rs, substrate.rs); microfacet is an rl_microfacet
int rlh_material_metal(const float specular[3], const float eta[3], const float k[3], uint32_t microfacet, float alpha, rl_material *out) {
    try {
        *out = Material::metal(Color{specular[0], specular[1], specular[2]}, Color{eta[0], eta[1], eta[2]}, Color{k[0], k[1], k[2]}, microfacet, alpha).m;
        return 0;
    } catch (const std::exception &) {
        return -1;
    }
}
int rlh_material_glass(const float reflectance[3], const float transmittance[3], float int_ior, float ext_ior, rl_material *out) {
    try {
        *out = Material::glass(Color{reflectance[0], reflectance[1], reflectance[2]}, Color{transmittance[0], transmittance[1], transmittance[2]}, int_ior, ext_ior).m;
        return 0;
    } catch (const std::exception &) {
        return -1;
    }
}
int rlh_material_substrate(const float diffuse[3], const float specular[3], uint32_t microfacet, float alpha, rl_material *out) {
    try {
        *out = Material::substrate(Color{diffuse[0], diffuse[1], diffuse[2]}, Color{specular[0], specular[1], specular[2]}, microfacet, alpha).m;
        return 0;
    } catch (const std::exception &) {
        return -1;
    }
}
float rlh_remap_roughness(float v, int remap) { return remap_roughness(v, remap != 0); }

// Returns the number of bytes needed (including the terminator); writes at most buflen.
size_t rlh_scene_to_json(const rlh_scene *s, char *buf, size_t buflen) {
    std::string j = scene_to_json(s->scene);
    if (buf && buflen) {
        size_t n = j.size() < buflen - 1 ? j.size() : buflen - 1;
        std::memcpy(buf, j.data(), n);
        buf[n] = 0;
    }
    return j.size() + 1;
}

// Camera::new, src/camera.rs:31-67.  fov_axis: 0 = Fov::Y, 1 = Fov::X.
int rlh_camera_create(uint32_t w, uint32_t h, int fov_axis, float fov_deg, const float to_world[16], int flip,
                      float out_sample_to_camera[16], float out_camera_to_sample[16]) {
    try {
        Mat4 m;
        std::memcpy(m.m, to_world, sizeof(m.m));
        Camera c = Camera::create(w, h, fov_axis ? Fov::X : Fov::Y, fov_deg, m, flip != 0);
        if (out_sample_to_camera) std::memcpy(out_sample_to_camera, c.sample_to_camera.m, sizeof(m.m));
        if (out_camera_to_sample) std::memcpy(out_camera_to_sample, c.camera_to_sample.m, sizeof(m.m));
        return 0;
    } catch (const std::exception &) {
        return -1;
    }
}

// Bitmap::save_pfm, src/structure.rs:547-560
int rlh_save_pfm(const char *path, uint32_t w, uint32_t h, const float *rgb) {
    try {
        Bitmap b;
        b.size_x = w, b.size_y = h;
        b.colors.assign(rgb, rgb + 3 * (size_t)w * h);
        b.save_pfm(path);
        return 0;
    } catch (const std::exception &) {
        return -1;
    }
}

int rlh_read_pfm(const char *path, uint32_t *w, uint32_t *h, float *rgb, size_t capacity_floats) {
    try {
        Bitmap b = Bitmap::read_pfm(path);
        *w = b.size_x, *h = b.size_y;
        if (rgb) {
            if (capacity_floats < b.colors.size()) return -2;
            std::memcpy(rgb, b.colors.data(), b.colors.size() * sizeof(float));
        }
        return 0;
    } catch (const std::exception &) {
        return -1;
    }
}

// Bitmap::save / Bitmap::read by extension (structure.rs:528-545, 670-683): .pfm, .png
int rlh_save_image(const char *path, uint32_t w, uint32_t h, const float *rgb) {
    try {
        Bitmap b;
        b.size_x = w, b.size_y = h;
        b.colors.assign(rgb, rgb + 3 * (size_t)w * h);
        b.save(path);
        return 0;
    } catch (const std::exception &) {
        return -1;
    }
}
int rlh_read_image(const char *path, uint32_t *w, uint32_t *h, float *rgb, size_t capacity_floats) {
    try {
        Bitmap b = Bitmap::read(path);
        *w = b.size_x, *h = b.size_y;
        if (rgb) {
            if (capacity_floats < b.colors.size()) return -2;
            std::memcpy(rgb, b.colors.data(), b.colors.size() * sizeof(float));
        }
        return 0;
    } catch (const std::exception &) {
        return -1;
    }
}

} // extern "C"
