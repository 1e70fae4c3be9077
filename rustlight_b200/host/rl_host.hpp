// rl_host.hpp -- host-side mirror of the rustlight types that sit above the hot path.
//
// The reference's host code is Rust (no toolchain in this image), so the layer above the C ABI
// (include/rl_b200.h) is restated in C++ with the reference's names and argument meaning:
//   Camera::new / Camera::generate      src/camera.rs:31-91
//   Mesh                                src/geometry.rs:107-182
//   Scene                               src/scene.rs:16-30
//   SceneLoaderManager / PBRTSceneLoader  src/scene_loader.rs:21-58, 77-315
//   IntegratorPathTracing / IntegratorDirect  src/integrators/explicit/path.rs:14-20, direct.rs:5-8
//   BufferCollection / Bitmap::save_pfm src/integrators/mod.rs:48-216, src/structure.rs:547-560
// It only prepares inputs (flat rl_scene_desc) and stores outputs; all rendering happens in the
// CUDA library behind rl_render().
#pragma once
#include <array>
#include <cstdint>
#include <map>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "rl_b200.h"

namespace rlh {

struct Vec3 {
    float x = 0, y = 0, z = 0;
};

// Column-major 4x4, m[4*col+row], same storage as cgmath::Matrix4<f32>.
struct Mat4 {
    float m[16];
    static Mat4 identity();
    static Mat4 from_nonuniform_scale(float x, float y, float z);
    static Mat4 from_translation(float x, float y, float z);
    // cgmath::perspective(fovy: Rad, aspect, near, far)
    static Mat4 perspective(float fovy_rad, float aspect, float near, float far);
    static Mat4 look_at_pbrt(Vec3 eye, Vec3 at, Vec3 up); // pbrt LookAt -> world_to_camera
    static Mat4 rotate_deg(float angle, Vec3 axis);
    Mat4 operator*(const Mat4 &rhs) const;
    std::optional<Mat4> invert() const; // cgmath SquareMatrix::invert (cofactors)
    Vec3 transform_point(Vec3 p) const; // homogeneous multiply then * (1/w)
    Vec3 transform_vector(Vec3 v) const;
    float &at(int col, int row) { return m[4 * col + row]; }
    float at(int col, int row) const { return m[4 * col + row]; }
};

struct Color {
    float r = 0, g = 0, b = 0;
    // Color::luminance, src/structure.rs:173-176
    float luminance() const { return r * 0.212671f + g * 0.715160f + b * 0.072169f; }
};

enum class Fov { Y, X }; // src/camera.rs:17-20

// src/camera.rs:5-15
struct Camera {
    uint32_t img_x = 0, img_y = 0;
    Mat4 camera_to_sample, sample_to_camera, to_world, to_local;
    // Camera::new(img, fov, mat, flip), src/camera.rs:31-67
    static Camera create(uint32_t w, uint32_t h, Fov axis, float fov_deg, const Mat4 &mat, bool flip);
    void scale_image(float s); // src/camera.rs:73-78 (truncating)
    // kept so that scale_image can rebuild the matrices like cli.rs:355-370 does
    Fov fov_axis = Fov::Y;
    float fov_deg = 0;
    bool flip = false;
};

// BSDF description; only what the hot path supports (src/bsdfs/{diffuse,phong}.rs).
struct Material {
    rl_material m{};
    std::vector<rl_material> subs; // BSDFBlend: {bsdf1, bsdf2} (m.blend_a / m.blend_b are filled in by Scene::desc)
    // BSDFBlend (bsdfs/blend.rs): weight * a + (1 - weight) * b; both parts rough, neither a blend
    static Material blend(const Material &a, const Material &b, float weight);
    static Material diffuse(Color kd);
    // weight_specular per src/bsdfs/mod.rs:518-523
    static Material phong(Color kd, Color ks, float exponent);
    // BSDFMetal (bsdfs/metal.rs); microfacet == RL_MICROFACET_NONE is the pure specular case ("mirror" in bsdf_pbrt,
    // bsdfs/mod.rs:349-357: specular = Kr, eta = 1, k = 0)
    static Material metal(Color specular, Color eta, Color k, uint32_t microfacet, float alpha);
    // BSDFGlass with .eta(int_ior, ext_ior) (bsdfs/glass.rs:43-48)
    static Material glass(Color reflectance, Color transmittance, float int_ior, float ext_ior);
    // BSDFSubstrate (bsdfs/substrate.rs)
    static Material substrate(Color diffuse, Color specular, uint32_t microfacet, float alpha);
};
// distribution_pbrt's roughness remapping (bsdfs/mod.rs:265-276)
float remap_roughness(float v, bool remap);

// BSDFColor other than Constant (bsdfs/mod.rs:11-30), usable on the diffuse-reflectance slot of a material
struct Texture {
    rl_texture t{};             // t.pixels points into `pixels`
    std::vector<float> pixels;  // bitmap: 3 * width * height (Bitmap.colors)
    static Texture bitmap(uint32_t w, uint32_t h, std::vector<float> rgb);
    static Texture bitmap_file(const std::string &filename); // .pfm (Bitmap::read_pfm), .png / .jpg (Bitmap::read_ldr_image) or binary .ppm (P6, /255 like read_ldr_image)
    static Texture checkerboard(Color c0, Color c1, float ox, float oy, float sx, float sy);
    static Texture grid(Color c0, Color c1, float line_width, float ox, float oy, float sx, float sy);
};

// src/geometry.rs:107-119
struct Mesh {
    std::string name;
    std::vector<float> vertices;   // 3*n
    std::vector<uint32_t> indices; // 3*t
    std::vector<float> normals;    // 3*n or empty (Option::None)
    std::vector<float> uv;         // 2*n or empty
    Material bsdf;
    bool is_light = false; // emission != EmissionType::Zero
    Color emission;
    // EmissionType::HSV / Texture (geometry.rs:99-104; what `-x hvs-light` / `-x texture-light` make of every mesh light, cli.rs:410-429):
    // emission_kind = RL_EMISSION_HSV | RL_EMISSION_TEXTURE, emission.r = scale, emission_texture = 1 + index into Scene::textures
    uint32_t emission_kind = 0; // 0: Zero / Color according to is_light
    uint32_t emission_texture = 0;
};

// src/scene.rs:16-30 (the fields this path consumes)
struct Scene {
    Camera camera;
    std::vector<std::shared_ptr<Mesh>> meshes;
    size_t nb_samples = 1;
    std::optional<size_t> nb_threads;
    std::string output_img_path = "out.pfm";
    bool has_volume = false, has_environment = false;
    Color environment; // EnvironmentLight{EnvironmentLightColor::Constant(environment)} when has_environment (scene_loader.rs:241-258)
    void set_environment(Color c) { has_environment = true, environment = c, environment_texture = 0; }
    bool use_ats = false; // Scene::build_emitters(build_ats) (scene.rs:53, 118-120): `-x ats`
    uint32_t environment_texture = 0; // EnvironmentLightColor::Texture: 1-based id of a bitmap texture (add_texture), 0 = constant
    void set_environment_texture(uint32_t id); // EnvironmentLightColor::new_texture(image) (scene_loader.rs:259-270)
    // Scene.emitters before build_emitters (EmittersState::Unbuild): PointEmitter / DirectionalLight (scene_loader.rs:207-240)
    std::vector<rl_light_desc> lights;
    std::vector<Texture> textures; // referenced by Material.m.kd_texture (1-based)
    uint32_t add_texture(Texture t) { // returns the 1-based id to store in rl_material.kd_texture
        textures.push_back(std::move(t));
        return (uint32_t)textures.size();
    }
    // `-x hvs-light` / `-x texture-light` (examples/cli.rs:410-429): every mesh light becomes EmissionType::HSV { scale } or
    // EmissionType::Texture { scale, img } with scale = luminance of its colour (a light that already is one of the two keeps scale 1);
    // tex_id = a bitmap texture id from add_texture (the reference reads "butterfly.jpg").
    void override_lights_hsv();
    void override_lights_texture(uint32_t tex_id);
    void add_point_light(Color intensity, float x, float y, float z);
    void add_directional_light(Color intensity, float dx, float dy, float dz); // direction = normalize(to - from)

    // Flatten to the C-ABI description.  The returned struct points into `this` and into
    // the scratch vector kept alive inside the Scene.
    const rl_scene_desc *desc();
    size_t nb_triangles() const;

  private:
    std::vector<rl_mesh_desc> mesh_descs_;
    std::vector<rl_texture> texture_descs_;
    std::vector<rl_material> submaterial_descs_;
    rl_scene_desc desc_{};
};

// src/scene_loader.rs:18-58.  `pbrt` is the reference's own route; `json` is a new, documented
// format (the reference has no JSON loader at this commit: SURVEY.md F3).
struct SceneLoader {
    virtual ~SceneLoader() = default;
    virtual Scene load(const std::string &filename, bool use_shading_normal) const = 0;
};
struct PBRTSceneLoader : SceneLoader {
    Scene load(const std::string &filename, bool use_shading_normal) const override;
    // base_dir: where `Shape "plymesh" "string filename"` paths are resolved (the scene file's directory)
    Scene load_string(const std::string &text, bool use_shading_normal, const std::string &base_dir = "") const;
};
struct JSONSceneLoader : SceneLoader {
    Scene load(const std::string &filename, bool use_shading_normal) const override;
    Scene load_string(const std::string &text, bool use_shading_normal, const std::string &base_dir = "") const;
};
// Mitsuba 0.x XML subset (scene_loader.rs:318-795, bsdfs/mod.rs:395-612): the reference's only route to BSDFPhong
struct MTSSceneLoader : SceneLoader {
    Scene load(const std::string &filename, bool use_shading_normal) const override;
    Scene load_string(const std::string &text, bool use_shading_normal, const std::string &base_dir = "") const;
};
struct SceneLoaderManager {
    std::map<std::string, std::shared_ptr<SceneLoader>> loader;
    SceneLoaderManager(); // registers "pbrt", "json" and "xml"
    Scene load(const std::string &filename, bool use_shading_normal) const;
};
std::string scene_to_json(const Scene &scene);

// src/structure.rs:383-560 (the part the path needs)
struct Bitmap {
    uint32_t size_x = 0, size_y = 0;
    std::vector<float> colors; // 3*size_x*size_y, row-major y*W+x
    void save_pfm(const std::string &path) const; // rows bottom-to-top, abs(), LE f32
    static Bitmap read_pfm(const std::string &path);
    void save_png(const std::string &path) const;    // Bitmap::save_ldr_image + Color::to_rgba: (min(c, 1)^(1/2.2) * 255) as u8
    static Bitmap read_png(const std::string &path); // Bitmap::read_ldr_image: to_rgb8() / 255
    static Bitmap read_tga(const std::string &path);  // Truevision TGA (true colour / grey / colour-mapped, raw or RLE)
    static Bitmap read_jpeg(const std::string &path); // Bitmap::read_ldr_image for .jpg / .jpeg (host/jpeg.cpp: baseline + progressive Huffman JPEG)
    void save(const std::string &path) const;        // by extension (structure.rs:528-545): pfm | png
    static Bitmap read(const std::string &path);     // by extension (structure.rs:670-683): pfm | png | jpg
};
// src/integrators/mod.rs:48-52; only the "primal" buffer exists on this path.
struct BufferCollection {
    std::map<std::string, Bitmap> values;
    void save(const std::string &name, const std::string &filename) const { values.at(name).save(filename); }
};

struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};

} // namespace rlh
