// scene.cpp -- Mat4 (cgmath conventions), Camera::new, Scene flattening, Bitmap IO.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iterator>
#include <algorithm>

#include "rl_host.hpp"
#include <cctype>
#include <zlib.h> // inflate / deflate / crc32 of the PNG container (the reference reads and writes PNG through the `image` crate)

namespace rlh {

// ------------------------------------------------------------------------------------------
// Mat4: cgmath::Matrix4<f32> conventions (column-major, column vectors).
// ------------------------------------------------------------------------------------------
Mat4 Mat4::identity() {
    Mat4 r{};
    for (int i = 0; i < 16; i++) r.m[i] = (i % 5 == 0) ? 1.0f : 0.0f;
    return r;
}
Mat4 Mat4::from_nonuniform_scale(float x, float y, float z) {
    Mat4 r = identity();
    r.at(0, 0) = x;
    r.at(1, 1) = y;
    r.at(2, 2) = z;
    return r;
}
Mat4 Mat4::from_translation(float x, float y, float z) {
    Mat4 r = identity();
    r.at(3, 0) = x;
    r.at(3, 1) = y;
    r.at(3, 2) = z;
    return r;
}
// cgmath 0.18 `perspective(fovy, aspect, near, far)` == PerspectiveFov::into::<Matrix4>():
//   f = cot(fovy/2); c0r0 = f/aspect; c1r1 = f; c2r2 = (far+near)/(near-far); c2r3 = -1;
//   c3r2 = 2*far*near/(near-far)
Mat4 Mat4::perspective(float fovy_rad, float aspect, float near, float far) {
    Mat4 r{};
    std::memset(r.m, 0, sizeof(r.m));
    float f = 1.0f / std::tan(fovy_rad / 2.0f);
    r.at(0, 0) = f / aspect;
    r.at(1, 1) = f;
    r.at(2, 2) = (far + near) / (near - far);
    r.at(2, 3) = -1.0f;
    r.at(3, 2) = (2.0f * far * near) / (near - far);
    return r;
}
Mat4 Mat4::operator*(const Mat4 &rhs) const {
    Mat4 r{};
    for (int c = 0; c < 4; c++)
        for (int row = 0; row < 4; row++) {
            // self[0]*v.x + self[1]*v.y + self[2]*v.z + self[3]*v.w, left to right
            float acc = at(0, row) * rhs.at(c, 0);
            acc = acc + at(1, row) * rhs.at(c, 1);
            acc = acc + at(2, row) * rhs.at(c, 2);
            acc = acc + at(3, row) * rhs.at(c, 3);
            r.at(c, row) = acc;
        }
    return r;
}
static float det3(float a, float b, float c, float d, float e, float f, float g, float h, float i) {
    // columns (a,b,c) (d,e,f) (g,h,i)
    return a * (e * i - h * f) - d * (b * i - h * c) + g * (b * f - e * c);
}
std::optional<Mat4> Mat4::invert() const {
    // cofactor inverse; cgmath's invert() has the same structure (determinant, transposed
    // cofactors times 1/det).  Host-only: last-bit differences are tolerated.
    float cof[16];
    for (int c = 0; c < 4; c++)
        for (int r = 0; r < 4; r++) {
            float s[9];
            int k = 0;
            for (int cc = 0; cc < 4; cc++) {
                if (cc == c) continue;
                for (int rr = 0; rr < 4; rr++) {
                    if (rr == r) continue;
                    s[k++] = at(cc, rr);
                }
            }
            float d = det3(s[0], s[1], s[2], s[3], s[4], s[5], s[6], s[7], s[8]);
            cof[4 * c + r] = ((c + r) & 1) ? -d : d;
        }
    float det = at(0, 0) * cof[0] + at(0, 1) * cof[1] + at(0, 2) * cof[2] + at(0, 3) * cof[3];
    if (det == 0.0f) return std::nullopt;
    float inv_det = 1.0f / det;
    Mat4 out{};
    for (int c = 0; c < 4; c++)
        for (int r = 0; r < 4; r++) out.at(c, r) = cof[4 * r + c] * inv_det; // adjugate = cof^T
    return out;
}
Vec3 Mat4::transform_point(Vec3 p) const {
    float h[4];
    for (int row = 0; row < 4; row++) {
        float acc = at(0, row) * p.x;
        acc = acc + at(1, row) * p.y;
        acc = acc + at(2, row) * p.z;
        acc = acc + at(3, row) * 1.0f;
        h[row] = acc;
    }
    float iw = 1.0f / h[3];
    return Vec3{h[0] * iw, h[1] * iw, h[2] * iw};
}
Vec3 Mat4::transform_vector(Vec3 v) const {
    float h[3];
    for (int row = 0; row < 3; row++) {
        float acc = at(0, row) * v.x;
        acc = acc + at(1, row) * v.y;
        acc = acc + at(2, row) * v.z;
        acc = acc + at(3, row) * 0.0f;
        h[row] = acc;
    }
    return Vec3{h[0], h[1], h[2]};
}
static Vec3 sub(Vec3 a, Vec3 b) { return Vec3{a.x - b.x, a.y - b.y, a.z - b.z}; }
static Vec3 cross(Vec3 a, Vec3 b) {
    return Vec3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
static Vec3 normalize(Vec3 a) {
    float l = std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z);
    return Vec3{a.x / l, a.y / l, a.z / l};
}
// pbrt-v3 LookAt: left-handed camera space, +z forward.  Returns world_to_camera.
Mat4 Mat4::look_at_pbrt(Vec3 eye, Vec3 at_, Vec3 up) {
    Vec3 dir = normalize(sub(at_, eye));
    Vec3 right = normalize(cross(normalize(up), dir));
    Vec3 new_up = cross(dir, right);
    Mat4 c2w = identity();
    c2w.at(0, 0) = right.x, c2w.at(0, 1) = right.y, c2w.at(0, 2) = right.z;
    c2w.at(1, 0) = new_up.x, c2w.at(1, 1) = new_up.y, c2w.at(1, 2) = new_up.z;
    c2w.at(2, 0) = dir.x, c2w.at(2, 1) = dir.y, c2w.at(2, 2) = dir.z;
    c2w.at(3, 0) = eye.x, c2w.at(3, 1) = eye.y, c2w.at(3, 2) = eye.z;
    return *c2w.invert();
}
Mat4 Mat4::rotate_deg(float angle, Vec3 axis) {
    Vec3 a = normalize(axis);
    float rad = angle * 3.14159265358979323846f / 180.0f;
    float s = std::sin(rad), c = std::cos(rad);
    Mat4 r = identity();
    r.at(0, 0) = a.x * a.x + (1 - a.x * a.x) * c;
    r.at(1, 0) = a.x * a.y * (1 - c) - a.z * s;
    r.at(2, 0) = a.x * a.z * (1 - c) + a.y * s;
    r.at(0, 1) = a.x * a.y * (1 - c) + a.z * s;
    r.at(1, 1) = a.y * a.y + (1 - a.y * a.y) * c;
    r.at(2, 1) = a.y * a.z * (1 - c) - a.x * s;
    r.at(0, 2) = a.x * a.z * (1 - c) - a.y * s;
    r.at(1, 2) = a.y * a.z * (1 - c) + a.x * s;
    r.at(2, 2) = a.z * a.z + (1 - a.z * a.z) * c;
    return r;
}

// ------------------------------------------------------------------------------------------
// Camera::new, src/camera.rs:31-67
// ------------------------------------------------------------------------------------------
Camera Camera::create(uint32_t w, uint32_t h, Fov axis, float fov_deg, const Mat4 &mat, bool flip) {
    if (w == 0 || h == 0) throw Error("Camera: empty image");
    Camera cam;
    cam.img_x = w;
    cam.img_y = h;
    cam.fov_axis = axis;
    cam.fov_deg = fov_deg;
    cam.flip = flip;
    cam.to_world = mat;
    auto inv = mat.invert();
    if (!inv) throw Error("Camera: to_world is singular");
    cam.to_local = *inv;
    float x_v = flip ? 1.0f : -1.0f;
    float aspect_ratio = (float)w / (float)h;
    const float PI = 3.14159265358979323846f;
    // Fov::Y(v) is converted as v*aspect (not through tan): src/camera.rs:41-44
    float fov_rad = (axis == Fov::X) ? fov_deg * PI / 180.0f : fov_deg * aspect_ratio * PI / 180.0f;
    cam.camera_to_sample = Mat4::from_nonuniform_scale(-0.5f, -0.5f * aspect_ratio, 1.0f) *
                           Mat4::from_translation(-1.0f, -1.0f / aspect_ratio, 0.0f) *
                           Mat4::perspective(fov_rad, 1.0f, 1e-2f, 1000.0f) *
                           Mat4::from_nonuniform_scale(x_v, 1.0f, -1.0f);
    auto s2c = cam.camera_to_sample.invert();
    if (!s2c) throw Error("Camera: camera_to_sample is singular");
    cam.sample_to_camera = *s2c;
    return cam;
}
void Camera::scale_image(float s) {
    uint32_t nx = (uint32_t)(s * (float)img_x), ny = (uint32_t)(s * (float)img_y);
    // the reference only rewrites `img` (camera.rs:73-78): matrices keep the original aspect
    img_x = nx;
    img_y = ny;
}

// ------------------------------------------------------------------------------------------
// Materials
// ------------------------------------------------------------------------------------------
Material Material::diffuse(Color kd) {
    Material r;
    r.m.kind = RL_BSDF_DIFFUSE;
    r.m.kd[0] = kd.r, r.m.kd[1] = kd.g, r.m.kd[2] = kd.b;
    r.m.ks[0] = r.m.ks[1] = r.m.ks[2] = 0.0f;
    r.m.exponent = 0.0f;
    r.m.weight_specular = 0.0f;
    return r;
}
Material Material::phong(Color kd, Color ks, float exponent) {
    Material r;
    r.m.kind = RL_BSDF_PHONG;
    r.m.kd[0] = kd.r, r.m.kd[1] = kd.g, r.m.kd[2] = kd.b;
    r.m.ks[0] = ks.r, r.m.ks[1] = ks.g, r.m.ks[2] = ks.b;
    r.m.exponent = exponent;
    float d_avg = kd.luminance(), s_avg = ks.luminance();
    if (d_avg + s_avg == 0.0f) throw Error("Phong: kd and ks are both black (bsdfs/mod.rs:521)");
    r.m.weight_specular = s_avg / (d_avg + s_avg);
    return r;
}

static void set3(float *dst, Color c) { dst[0] = c.r, dst[1] = c.g, dst[2] = c.b; }
Material Material::metal(Color specular, Color eta, Color k, uint32_t microfacet, float alpha) {
    if (microfacet > RL_MICROFACET_BECKMANN) throw Error("metal: unknown microfacet distribution");
    Material r;
    r.m.kind = RL_BSDF_METAL;
    set3(r.m.ks, specular), set3(r.m.eta, eta), set3(r.m.k, k);
    r.m.microfacet = microfacet;
    r.m.alpha = microfacet == RL_MICROFACET_NONE ? 0.0f : alpha;
    return r;
}
Material Material::glass(Color reflectance, Color transmittance, float int_ior, float ext_ior) {
    Material r;
    r.m.kind = RL_BSDF_GLASS;
    set3(r.m.ks, reflectance), set3(r.m.kt, transmittance);
    r.m.ior = int_ior / ext_ior;
    if (r.m.ior == 0.0f) throw Error("glass: eta must not be 0 (bsdfs/glass.rs:45)");
    return r;
}
Material Material::substrate(Color diffuse, Color specular, uint32_t microfacet, float alpha) {
    if (microfacet > RL_MICROFACET_BECKMANN) throw Error("substrate: unknown microfacet distribution");
    Material r;
    r.m.kind = RL_BSDF_SUBSTRATE;
    set3(r.m.kd, diffuse), set3(r.m.ks, specular);
    r.m.microfacet = microfacet;
    r.m.alpha = microfacet == RL_MICROFACET_NONE ? 0.0f : alpha;
    return r;
}
static bool blend_part_ok(const rl_material &m) { // blend.rs:17 asserts !is_smooth() on both parts; here also: no nested blend, no textures
    const bool rough = m.kind == RL_BSDF_DIFFUSE || m.kind == RL_BSDF_PHONG ||
                       ((m.kind == RL_BSDF_METAL || m.kind == RL_BSDF_SUBSTRATE) && m.microfacet != RL_MICROFACET_NONE);
    return rough && !m.kd_texture && !m.ks_texture && !m.kt_texture && !m.eta_texture && !m.k_texture;
}
Material Material::blend(const Material &a, const Material &b, float weight) {
    if (!blend_part_ok(a.m) || !blend_part_ok(b.m)) throw Error("blend: both parts must be rough BSDFs with constant colours (bsdfs/blend.rs:17)");
    if (!(weight >= 0.0f && weight <= 1.0f)) throw Error("blend: weight outside [0, 1]");
    Material r;
    r.m.kind = RL_BSDF_BLEND;
    r.m.blend_weight = weight;
    r.subs = {a.m, b.m};
    return r;
}
float remap_roughness(float v, bool remap) {
    if (!remap) return v;
    float x = std::log(std::max(v, 1e-3f));
    return 1.62142f + 0.819955f * x + 0.1734f * x * x + 0.0171201f * x * x * x + 0.000640711f * x * x * x * x;
}

// ------------------------------------------------------------------------------------------
// Scene
// ------------------------------------------------------------------------------------------
size_t Scene::nb_triangles() const {
    size_t n = 0;
    for (auto &m : meshes) n += m->indices.size() / 3;
    return n;
}
Texture Texture::bitmap(uint32_t w, uint32_t h, std::vector<float> rgb) {
    if (w == 0 || h == 0 || rgb.size() != (size_t)3 * w * h) throw Error("texture: bitmap size mismatch");
    Texture t;
    t.t.kind = RL_TEX_BITMAP, t.t.width = w, t.t.height = h;
    t.pixels = std::move(rgb);
    return t;
}
Texture Texture::bitmap_file(const std::string &filename) {
    auto ends = [&](const char *e) { return filename.size() >= std::strlen(e) && filename.compare(filename.size() - std::strlen(e), std::string::npos, e) == 0; };
    if (ends(".pfm")) {
        Bitmap b = Bitmap::read_pfm(filename);
        return bitmap(b.size_x, b.size_y, std::move(b.colors));
    }
    if (ends(".ppm")) { // binary P6, 8 bit: value / 255 (Bitmap::read_ldr_image, structure.rs:649-668)
        std::ifstream f(filename, std::ios::binary);
        if (!f) throw Error("cannot open " + filename);
        std::string magic;
        f >> magic;
        auto next_int = [&]() {
            for (;;) {
                int c = f.peek();
                if (c == '#') {
                    std::string skip;
                    std::getline(f, skip);
                } else if (std::isspace(c)) f.get();
                else break;
            }
            long v;
            f >> v;
            return v;
        };
        long w = next_int(), h = next_int(), maxv = next_int();
        f.get();
        if (magic != "P6" || w <= 0 || h <= 0 || maxv != 255) throw Error("ppm: only binary P6 with maxval 255: " + filename);
        std::vector<unsigned char> raw((size_t)3 * w * h);
        f.read(reinterpret_cast<char *>(raw.data()), (std::streamsize)raw.size());
        if (!f) throw Error("ppm: truncated " + filename);
        std::vector<float> rgb(raw.size());
        for (size_t i = 0; i < raw.size(); i++) rgb[i] = (float)raw[i] / 255.0f;
        return bitmap((uint32_t)w, (uint32_t)h, std::move(rgb));
    }
    if (ends(".png")) { // Bitmap::read -> read_ldr_image (structure.rs:649-668): image::open(..).to_rgb8(), value / 255, no gamma
        Bitmap b = Bitmap::read_png(filename);
        return bitmap(b.size_x, b.size_y, std::move(b.colors));
    }
    if (ends(".jpg") || ends(".jpeg") || ends(".JPG")) {
        Bitmap b = Bitmap::read_jpeg(filename);
        return bitmap(b.size_x, b.size_y, std::move(b.colors));
    }
    if (ends(".tga") || ends(".TGA")) {
        Bitmap b = Bitmap::read_tga(filename);
        return bitmap(b.size_x, b.size_y, std::move(b.colors));
    }
    throw Error("texture: only .pfm, .png, .jpg, .tga and .ppm images can be read here: " + filename);
}
Texture Texture::checkerboard(Color c0, Color c1, float ox, float oy, float sx, float sy) {
    Texture t;
    t.t.kind = RL_TEX_CHECKERBOARD;
    set3(t.t.color0, c0), set3(t.t.color1, c1);
    t.t.offset[0] = ox, t.t.offset[1] = oy, t.t.scale[0] = sx, t.t.scale[1] = sy;
    return t;
}
Texture Texture::grid(Color c0, Color c1, float line_width, float ox, float oy, float sx, float sy) {
    Texture t = checkerboard(c0, c1, ox, oy, sx, sy);
    t.t.kind = RL_TEX_GRID;
    t.t.line_width = line_width;
    return t;
}
void Scene::add_point_light(Color intensity, float x, float y, float z) {
    rl_light_desc l{};
    l.kind = RL_LIGHT_POINT;
    l.intensity[0] = intensity.r, l.intensity[1] = intensity.g, l.intensity[2] = intensity.b;
    l.v[0] = x, l.v[1] = y, l.v[2] = z;
    lights.push_back(l);
}
void Scene::add_directional_light(Color intensity, float dx, float dy, float dz) {
    float len = std::sqrt(dx * dx + dy * dy + dz * dz);
    if (!(len > 0.0f)) throw Error("directional light: zero direction");
    float inv = 1.0f / len; // cgmath normalize: v * (1 / |v|)
    rl_light_desc l{};
    l.kind = RL_LIGHT_DIRECTIONAL;
    l.intensity[0] = intensity.r, l.intensity[1] = intensity.g, l.intensity[2] = intensity.b;
    l.v[0] = dx * inv, l.v[1] = dy * inv, l.v[2] = dz * inv;
    lights.push_back(l);
}
void Scene::set_environment_texture(uint32_t id) {
    if (id == 0 || id > textures.size() || textures[id - 1].t.kind != RL_TEX_BITMAP) throw Error("environment texture: not a bitmap texture id");
    has_environment = true, environment = Color{0.0f, 0.0f, 0.0f}, environment_texture = id;
}
// examples/cli.rs:410-429 (Color::luminance, structure.rs:173-176)
static float light_scale(const Mesh &m) {
    if (m.emission_kind == RL_EMISSION_HSV || m.emission_kind == RL_EMISSION_TEXTURE) return 1.0f; // `_ => 1.0`
    return m.emission.r * 0.212671f + m.emission.g * 0.715160f + m.emission.b * 0.072169f;
}
void Scene::override_lights_hsv() {
    for (auto &mp : meshes) {
        Mesh &m = *mp;
        if (!m.is_light) continue;
        const float scale = light_scale(m);
        m.emission_kind = RL_EMISSION_HSV, m.emission = Color{scale, scale, scale}, m.emission_texture = 0;
    }
}
void Scene::override_lights_texture(uint32_t tex_id) {
    if (tex_id == 0 || tex_id > textures.size() || textures[tex_id - 1].t.kind != RL_TEX_BITMAP) throw Error("texture lights: not a bitmap texture id");
    for (auto &mp : meshes) {
        Mesh &m = *mp;
        if (!m.is_light) continue;
        const float scale = light_scale(m);
        m.emission_kind = RL_EMISSION_TEXTURE, m.emission = Color{scale, scale, scale}, m.emission_texture = tex_id;
    }
}
const rl_scene_desc *Scene::desc() {
    mesh_descs_.clear();
    submaterial_descs_.clear();
    for (auto &mp : meshes) {
        const Mesh &m = *mp;
        rl_mesh_desc d{};
        d.P = m.vertices.data();
        d.nverts = (uint32_t)(m.vertices.size() / 3);
        d.idx = m.indices.data();
        d.ntris = (uint32_t)(m.indices.size() / 3);
        d.N = m.normals.empty() ? nullptr : m.normals.data();
        d.UV = m.uv.empty() ? nullptr : m.uv.data();
        d.mat = m.bsdf.m;
        if (d.mat.kind == RL_BSDF_BLEND && m.bsdf.subs.size() == 2) {
            submaterial_descs_.push_back(m.bsdf.subs[0]);
            submaterial_descs_.push_back(m.bsdf.subs[1]);
            d.mat.blend_a = (uint32_t)submaterial_descs_.size() - 1u, d.mat.blend_b = (uint32_t)submaterial_descs_.size();
        }
        d.emission_kind = m.is_light ? (m.emission_kind ? m.emission_kind : 1u) : 0u;
        d.emission[0] = m.emission.r, d.emission[1] = m.emission.g, d.emission[2] = m.emission.b;
        d.emission_texture = m.emission_texture;
        mesh_descs_.push_back(d);
    }
    desc_.nmeshes = (uint32_t)mesh_descs_.size();
    desc_.meshes = mesh_descs_.data();
    desc_.camera.width = camera.img_x;
    desc_.camera.height = camera.img_y;
    std::memcpy(desc_.camera.sample_to_camera, camera.sample_to_camera.m, sizeof(float) * 16);
    std::memcpy(desc_.camera.to_world, camera.to_world.m, sizeof(float) * 16);
    desc_.has_volume = has_volume ? 1u : 0u;
    desc_.has_environment = has_environment ? (environment_texture ? 2u : 1u) : 0u;
    desc_.environment_texture = environment_texture;
    desc_.use_ats = use_ats ? 1u : 0u;
    desc_.environment[0] = environment.r, desc_.environment[1] = environment.g, desc_.environment[2] = environment.b;
    texture_descs_.clear();
    for (auto &t : textures) {
        rl_texture d = t.t;
        d.pixels = t.pixels.empty() ? nullptr : t.pixels.data();
        texture_descs_.push_back(d);
    }
    desc_.ntextures = (uint32_t)texture_descs_.size();
    desc_.textures = texture_descs_.empty() ? nullptr : texture_descs_.data();
    desc_.nsubmaterials = (uint32_t)submaterial_descs_.size();
    desc_.submaterials = submaterial_descs_.empty() ? nullptr : submaterial_descs_.data();
    desc_.nlights = (uint32_t)lights.size();
    desc_.lights = lights.empty() ? nullptr : lights.data();
    return &desc_;
}

// ------------------------------------------------------------------------------------------
// Bitmap::save_pfm, src/structure.rs:547-560
// ------------------------------------------------------------------------------------------
void Bitmap::save_pfm(const std::string &path) const {
    std::ofstream f(path, std::ios::binary);
    if (!f) throw Error("cannot open " + path);
    char header[64];
    int n = std::snprintf(header, sizeof(header), "PF\n%u %u\n-1.0\n", size_x, size_y);
    f.write(header, n);
    std::vector<float> row(3 * (size_t)size_x);
    for (uint32_t y = 0; y < size_y; y++) {
        const float *src = &colors[3 * (size_t)(size_y - y - 1) * size_x];
        for (size_t i = 0; i < row.size(); i++) row[i] = std::fabs(src[i]);
        f.write(reinterpret_cast<const char *>(row.data()), (std::streamsize)(row.size() * sizeof(float)));
    }
}
Bitmap Bitmap::read_pfm(const std::string &path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw Error("cannot open " + path);
    std::string magic;
    uint32_t w, h;
    float scale;
    f >> magic >> w >> h >> scale;
    f.get(); // single whitespace after the scale
    if (magic != "PF") throw Error("not a colour PFM: " + path);
    Bitmap b;
    b.size_x = w, b.size_y = h;
    b.colors.resize(3 * (size_t)w * h);
    for (uint32_t y = 0; y < h; y++)
        f.read(reinterpret_cast<char *>(&b.colors[3 * (size_t)(h - y - 1) * w]), (std::streamsize)(3 * w * sizeof(float)));
    return b;
}

// ------------------------------------------------------------------------------------------
// PNG (the `image` feature of the reference, on by default: Bitmap::read_ldr_image / save_ldr_image, structure.rs:471-485, 649-668)
// ------------------------------------------------------------------------------------------
// Reader: what `image::open(path).to_rgb8()` yields for a PNG, then value / 255 (no gamma, alpha dropped, gAMA / tRNS ignored):
// colour types 0 / 2 / 3 / 4 / 6, bit depths 1-16, the five row filters; sub-byte grey samples are scaled to 8 bits (x 255 / 85 / 17) and
// 16-bit samples narrowed as image 0.24 does, (v + 128) / 257.  Interlaced (Adam7) files are refused loudly.
namespace {
uint32_t be32(const unsigned char *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3]; }
void put_be32(std::vector<unsigned char> &v, uint32_t x) {
    v.push_back((unsigned char)(x >> 24)), v.push_back((unsigned char)(x >> 16)), v.push_back((unsigned char)(x >> 8)), v.push_back((unsigned char)x);
}
int paeth(int a, int b, int c) {
    const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}
} // namespace
Bitmap Bitmap::read_png(const std::string &path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw Error("cannot open " + path);
    std::vector<unsigned char> file((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (file.size() < 8 || std::memcmp(file.data(), sig, 8) != 0) throw Error("not a PNG file: " + path);
    uint32_t w = 0, h = 0;
    int depth = 0, ctype = -1;
    std::vector<unsigned char> idat, plte;
    bool end = false;
    for (size_t pos = 8; !end;) {
        if (pos + 12 > file.size()) throw Error("png: truncated " + path);
        const uint32_t len = be32(&file[pos]);
        if ((size_t)len > file.size() - pos - 12) throw Error("png: truncated chunk in " + path);
        const unsigned char *type = &file[pos + 4], *data = &file[pos + 8];
        if (be32(data + len) != (uint32_t)crc32(crc32(0L, type, 4), data, len)) throw Error("png: CRC mismatch in " + path);
        if (std::memcmp(type, "IHDR", 4) == 0) {
            if (len != 13) throw Error("png: bad IHDR in " + path);
            w = be32(data), h = be32(data + 4), depth = data[8], ctype = data[9];
            if (data[10] != 0 || data[11] != 0) throw Error("png: unknown compression / filter method in " + path);
            if (data[12] != 0) throw Error("png: interlaced (Adam7) files are not supported: " + path);
        } else if (std::memcmp(type, "PLTE", 4) == 0) plte.assign(data, data + len);
        else if (std::memcmp(type, "IDAT", 4) == 0) idat.insert(idat.end(), data, data + len);
        else if (std::memcmp(type, "IEND", 4) == 0) end = true;
        pos += (size_t)len + 12;
    }
    int channels;
    switch (ctype) {
    case 0: channels = 1; break;
    case 2: channels = 3; break;
    case 3: channels = 1; break;
    case 4: channels = 2; break;
    case 6: channels = 4; break;
    default: throw Error("png: bad colour type in " + path);
    }
    const bool depth_ok = ctype == 0 ? (depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16)
                          : ctype == 3 ? (depth == 1 || depth == 2 || depth == 4 || depth == 8) : (depth == 8 || depth == 16);
    if (w == 0 || h == 0 || !depth_ok || (uint64_t)w * h > (1ull << 28)) throw Error("png: bad header in " + path);
    const size_t row_bytes = ((size_t)w * channels * depth + 7) / 8, bpp = std::max<size_t>(1, (size_t)channels * depth / 8);
    std::vector<unsigned char> raw((row_bytes + 1) * h);
    uLongf got = (uLongf)raw.size();
    if (uncompress(raw.data(), &got, idat.data(), (uLong)idat.size()) != Z_OK || got != raw.size()) throw Error("png: corrupt image data in " + path);
    std::vector<unsigned char> prev(row_bytes, 0), cur(row_bytes);
    Bitmap b;
    b.size_x = w, b.size_y = h;
    b.colors.resize((size_t)3 * w * h);
    auto narrow16 = [](uint32_t v) { return (v + 128u) / 257u; }; // image 0.24: u16 -> u8
    for (uint32_t y = 0; y < h; y++) {
        const unsigned char *src = &raw[(row_bytes + 1) * y];
        const int ft = src[0];
        if (ft > 4) throw Error("png: bad row filter in " + path);
        for (size_t i = 0; i < row_bytes; i++) {
            const int a = i >= bpp ? cur[i - bpp] : 0, up = prev[i], c = i >= bpp ? prev[i - bpp] : 0;
            const int pred = ft == 0 ? 0 : ft == 1 ? a : ft == 2 ? up : ft == 3 ? (a + up) / 2 : paeth(a, up, c);
            cur[i] = (unsigned char)(src[1 + i] + pred);
        }
        for (uint32_t x = 0; x < w; x++) {
            uint32_t px[4] = {0, 0, 0, 0};
            for (int c = 0; c < channels; c++) {
                if (depth == 8) px[c] = cur[(size_t)x * channels + c];
                else if (depth == 16) px[c] = ((uint32_t)cur[((size_t)x * channels + c) * 2] << 8) | cur[((size_t)x * channels + c) * 2 + 1];
                else { // 1 / 2 / 4 bits, one channel, most significant bits first
                    const size_t bit = (size_t)x * depth;
                    px[c] = (cur[bit / 8] >> (8 - depth - (bit % 8))) & ((1u << depth) - 1u);
                }
            }
            uint32_t r, g, bl;
            if (ctype == 3) {
                if ((size_t)px[0] * 3 + 2 >= plte.size()) throw Error("png: palette index out of range in " + path);
                r = plte[px[0] * 3], g = plte[px[0] * 3 + 1], bl = plte[px[0] * 3 + 2];
            } else {
                for (int c = 0; c < channels; c++) {
                    if (depth == 16) px[c] = narrow16(px[c]);
                    else if (depth < 8) px[c] *= 255u / ((1u << depth) - 1u);
                }
                if (channels <= 2) r = g = bl = px[0];
                else r = px[0], g = px[1], bl = px[2];
            }
            float *dst = &b.colors[3 * ((size_t)y * w + x)];
            dst[0] = (float)r / 255.0f, dst[1] = (float)g / 255.0f, dst[2] = (float)bl / 255.0f;
        }
        prev.swap(cur);
    }
    return b;
}
// Writer: Bitmap::save_ldr_image with Color::to_rgba (structure.rs:160-167, 471-485): (min(c, 1) ^ (1 / 2.2) * 255) as u8 -- truncation,
// f32::min semantics (a NaN channel becomes 1), Rust's saturating cast (the NaN of a negative base -> 0) -- as 8-bit RGB.  The bytes of the file differ from the `image` crate's encoder (filter
// and deflate choices), the decoded pixels do not.
void Bitmap::save_png(const std::string &path) const {
    auto to_u8 = [](float c) -> unsigned char {
        const float v = std::pow(std::fmin(c, 1.0f), 1.0f / 2.2f) * 255.0f; // fmin: f32::min returns the non-NaN operand
        if (!(v > 0.0f)) return 0;                                          // NaN (negative base) and 0
        return v >= 255.0f ? 255 : (unsigned char)v;
    };
    std::vector<unsigned char> raw(((size_t)3 * size_x + 1) * size_y);
    for (uint32_t y = 0; y < size_y; y++) {
        unsigned char *row = &raw[((size_t)3 * size_x + 1) * y];
        row[0] = 0; // filter: none
        for (size_t i = 0; i < (size_t)3 * size_x; i++) row[1 + i] = to_u8(colors[(size_t)3 * size_x * y + i]);
    }
    uLongf zlen = compressBound((uLong)raw.size());
    std::vector<unsigned char> z(zlen);
    if (compress2(z.data(), &zlen, raw.data(), (uLong)raw.size(), 6) != Z_OK) throw Error("png: deflate failed for " + path);
    std::vector<unsigned char> out = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    auto chunk = [&](const char *type, const unsigned char *data, size_t len) {
        put_be32(out, (uint32_t)len);
        const size_t at = out.size();
        out.insert(out.end(), type, type + 4);
        if (len) out.insert(out.end(), data, data + len);
        put_be32(out, (uint32_t)crc32(0L, &out[at], (uInt)(len + 4)));
    };
    std::vector<unsigned char> ihdr;
    put_be32(ihdr, size_x), put_be32(ihdr, size_y);
    ihdr.insert(ihdr.end(), {8, 2, 0, 0, 0}); // 8 bits, RGB, deflate, adaptive filtering, no interlace
    chunk("IHDR", ihdr.data(), ihdr.size());
    chunk("IDAT", z.data(), zlen);
    chunk("IEND", nullptr, 0);
    std::ofstream f(path, std::ios::binary);
    if (!f) throw Error("cannot open " + path);
    f.write(reinterpret_cast<const char *>(out.data()), (std::streamsize)out.size());
    if (!f) throw Error("cannot write " + path);
}
// Truevision TGA (textures of the pbrt-v3 scenes; read by the `image` crate in the reference): types 2 / 10 (true colour, raw / RLE, 24 or 32 bits: BGR(A),
// alpha dropped), 3 / 11 (grey 8 bits), 1 / 9 (colour-mapped 8-bit indices over a 24 / 32-bit map); bottom-up rows unless bit 5 of the descriptor says
// top-down, right-to-left columns when bit 4 is set.  value / 255 like every LDR image (structure.rs:649-668).
Bitmap Bitmap::read_tga(const std::string &path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw Error("cannot open " + path);
    std::vector<unsigned char> d((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    if (d.size() < 18) throw Error("tga: truncated header in " + path);
    const int id_len = d[0], cmap_type = d[1], type = d[2], cmap_first = d[3] | (d[4] << 8), cmap_len = d[5] | (d[6] << 8), cmap_bits = d[7];
    const int w = d[12] | (d[13] << 8), h = d[14] | (d[15] << 8), bpp = d[16], desc = d[17];
    const bool rle = type >= 9;
    const int base = type & 7;
    if ((base != 1 && base != 2 && base != 3) || w <= 0 || h <= 0) throw Error("tga: unsupported image type in " + path);
    if ((base == 2 && bpp != 24 && bpp != 32) || (base == 3 && bpp != 8) || (base == 1 && (bpp != 8 || cmap_type != 1 || (cmap_bits != 24 && cmap_bits != 32))))
        throw Error("tga: unsupported pixel format in " + path);
    size_t pos = 18 + (size_t)id_len;
    const size_t cmap_bytes = cmap_type ? (size_t)cmap_len * ((cmap_bits + 7) / 8) : 0;
    if (pos + cmap_bytes > d.size()) throw Error("tga: truncated colour map in " + path);
    const unsigned char *cmap = &d[pos];
    pos += cmap_bytes;
    const int px = bpp / 8;
    std::vector<unsigned char> raw((size_t)w * h * px);
    if (!rle) {
        if (pos + raw.size() > d.size()) throw Error("tga: truncated pixel data in " + path);
        std::memcpy(raw.data(), &d[pos], raw.size());
    } else {
        size_t o = 0;
        while (o < raw.size()) {
            if (pos >= d.size()) throw Error("tga: truncated RLE data in " + path);
            const int hd = d[pos++], n = (hd & 127) + 1;
            if (o + (size_t)n * px > raw.size()) throw Error("tga: RLE packet past the end of the image in " + path);
            if (hd & 128) {
                if (pos + px > d.size()) throw Error("tga: truncated RLE data in " + path);
                for (int k = 0; k < n; k++, o += px) std::memcpy(&raw[o], &d[pos], px);
                pos += px;
            } else {
                if (pos + (size_t)n * px > d.size()) throw Error("tga: truncated RLE data in " + path);
                std::memcpy(&raw[o], &d[pos], (size_t)n * px);
                o += (size_t)n * px, pos += (size_t)n * px;
            }
        }
    }
    Bitmap b;
    b.size_x = (uint32_t)w, b.size_y = (uint32_t)h;
    b.colors.resize((size_t)3 * w * h);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const int sy = (desc & 0x20) ? y : h - 1 - y, sx = (desc & 0x10) ? w - 1 - x : x;
            const unsigned char *p = &raw[((size_t)sy * w + sx) * px];
            unsigned char r, g, bl;
            if (base == 3) r = g = bl = p[0];
            else if (base == 2) bl = p[0], g = p[1], r = p[2];
            else {
                const int idx = (int)p[0] - cmap_first, eb = (cmap_bits + 7) / 8;
                if (idx < 0 || idx >= cmap_len) throw Error("tga: colour index out of range in " + path);
                bl = cmap[idx * eb], g = cmap[idx * eb + 1], r = cmap[idx * eb + 2];
            }
            float *dst = &b.colors[3 * ((size_t)y * w + x)];
            dst[0] = (float)r / 255.0f, dst[1] = (float)g / 255.0f, dst[2] = (float)bl / 255.0f;
        }
    return b;
}
// Bitmap::save / Bitmap::read (structure.rs:528-545, 670-683): by extension.  .exr needs the reference's optional `openexr` feature
// (off by default: it panics there) and is refused here.
static std::string extension_of(const std::string &path) {
    const size_t dot = path.find_last_of('.'), slash = path.find_last_of('/');
    if (dot == std::string::npos || (slash != std::string::npos && dot < slash)) throw Error("No file extension provided: " + path);
    return path.substr(dot + 1);
}
void Bitmap::save(const std::string &path) const {
    const std::string ext = extension_of(path);
    if (ext == "pfm") save_pfm(path);
    else if (ext == "png") save_png(path);
    else if (ext == "exr") throw Error("OpenEXR output is not built (the reference's optional `openexr` feature): " + path);
    else throw Error("Unknow output file extension: " + path);
}
Bitmap Bitmap::read(const std::string &path) {
    const std::string ext = extension_of(path);
    if (ext == "pfm") return read_pfm(path);
    if (ext == "png") return read_png(path);
    if (ext == "jpg" || ext == "jpeg" || ext == "JPG") return read_jpeg(path);
    if (ext == "tga" || ext == "TGA") return read_tga(path);
    throw Error("image: only .pfm, .png, .jpg and .tga can be read: " + path);
}

} // namespace rlh
