// loaders.cpp -- scene loaders above the hot path.
//
// PBRTSceneLoader mirrors src/scene_loader.rs:77-315 for the pbrt-v3 subset the path's scenes
// use: Transform/ConcatTransform/LookAt/Translate/Scale/Rotate/Identity, Film, Camera
// "perspective", MakeNamedMaterial/NamedMaterial/Material "matte" | "mirror" | "metal" | "glass" | "substrate",
// Shape "trianglemesh" | "plymesh", ObjectBegin/ObjectEnd/ObjectInstance, AreaLightSource "diffuse",
// LightSource "point" | "distant", Texture "imagemap", Include, AttributeBegin/End, TransformBegin/End, ReverseOrientation.
// The pbrt_rs crate that does the parsing for the reference is not vendored,
// so the grammar follows the pbrt-v3 file format itself.
// One labelled extension: material type "phong" (Kd, Ks, exponent) -> BSDFPhong, which the
// reference can only build through its Mitsuba route (src/bsdfs/mod.rs:509-531).
//
// JSONSceneLoader reads the new format documented in DESIGN.md ("scene JSON"); the reference
// has no JSON loader at this commit.
#include <algorithm>
#include <array>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <zlib.h> // Mitsuba "serialized" meshes are deflate streams

#include "rl_host.hpp"

namespace rlh {

static std::string read_file(const std::string &filename) {
    std::ifstream f(filename, std::ios::binary);
    if (!f) throw Error("cannot open scene file: " + filename);
    std::stringstream ss;
    ss << f.rdbuf();
    return ss.str();
}

// ------------------------------------------------------------------------------------------
// pbrt tokenizer
// ------------------------------------------------------------------------------------------
namespace {
struct Tok {
    enum Kind { Ident, Str, Num, LBr, RBr, End } kind = End;
    std::string s;
    double v = 0;
};
struct Lexer {
    const std::string &t;
    size_t p = 0;
    explicit Lexer(const std::string &text) : t(text) {}
    Tok next() {
        for (;;) {
            while (p < t.size() && std::isspace((unsigned char)t[p])) p++;
            if (p < t.size() && t[p] == '#') {
                while (p < t.size() && t[p] != '\n') p++;
                continue;
            }
            break;
        }
        Tok k;
        if (p >= t.size()) return k;
        char c = t[p];
        if (c == '[') { p++; k.kind = Tok::LBr; return k; }
        if (c == ']') { p++; k.kind = Tok::RBr; return k; }
        if (c == '"') {
            size_t e = t.find('"', p + 1);
            if (e == std::string::npos) throw Error("pbrt: unterminated string");
            k.kind = Tok::Str;
            k.s = t.substr(p + 1, e - p - 1);
            p = e + 1;
            return k;
        }
        if (std::isdigit((unsigned char)c) || c == '-' || c == '+' || c == '.') {
            char *end = nullptr;
            k.v = std::strtod(t.c_str() + p, &end);
            if (end == t.c_str() + p) throw Error("pbrt: bad number");
            k.kind = Tok::Num;
            p = (size_t)(end - t.c_str());
            return k;
        }
        size_t e = p;
        while (e < t.size() && (std::isalnum((unsigned char)t[e]) || t[e] == '_')) e++;
        if (e == p) throw Error(std::string("pbrt: unexpected character '") + c + "'");
        k.kind = Tok::Ident;
        k.s = t.substr(p, e - p);
        p = e;
        return k;
    }
};

struct Param {
    std::string type, name;
    std::vector<double> nums;
    std::vector<std::string> strs;
};
struct ParamSet {
    std::vector<Param> ps;
    const Param *find(const std::string &name) const {
        for (auto &p : ps)
            if (p.name == name) return &p;
        return nullptr;
    }
    std::string str(const std::string &name, const std::string &def = "") const {
        auto p = find(name);
        return (p && !p->strs.empty()) ? p->strs[0] : def;
    }
    double num(const std::string &name, double def) const {
        auto p = find(name);
        return (p && !p->nums.empty()) ? p->nums[0] : def;
    }
    bool boolean(const std::string &name, bool def) const {
        auto p = find(name);
        if (!p || p->strs.empty()) return def;
        return p->strs[0] == "true";
    }
    Color rgb(const std::string &name, Color def) const {
        auto p = find(name);
        if (!p) return def;
        if (p->type == "texture" || p->type == "spectrum" || p->type == "blackbody")
            throw Error("pbrt: '" + p->type + " " + name + "' is not supported on this path");
        if (p->nums.size() == 1) return Color{(float)p->nums[0], (float)p->nums[0], (float)p->nums[0]};
        if (p->nums.size() < 3) throw Error("pbrt: rgb parameter needs 3 values: " + name);
        return Color{(float)p->nums[0], (float)p->nums[1], (float)p->nums[2]};
    }
};

struct Parser {
    Lexer lex;
    Tok cur;
    explicit Parser(const std::string &text) : lex(text) { cur = lex.next(); }
    void advance() { cur = lex.next(); }
    std::string expect_str() {
        if (cur.kind != Tok::Str) throw Error("pbrt: expected a quoted string");
        std::string s = cur.s;
        advance();
        return s;
    }
    double expect_num() {
        if (cur.kind != Tok::Num) throw Error("pbrt: expected a number");
        double v = cur.v;
        advance();
        return v;
    }
    std::vector<double> nums(size_t n) {
        std::vector<double> r;
        bool br = cur.kind == Tok::LBr;
        if (br) advance();
        for (size_t i = 0; i < n; i++) r.push_back(expect_num());
        if (br) {
            if (cur.kind != Tok::RBr) throw Error("pbrt: expected ']'");
            advance();
        }
        return r;
    }
    ParamSet params() {
        ParamSet set;
        while (cur.kind == Tok::Str) {
            Param p;
            std::string decl = cur.s;
            advance();
            std::istringstream ds(decl);
            ds >> p.type >> p.name;
            if (p.name.empty()) throw Error("pbrt: bad parameter declaration \"" + decl + "\"");
            auto take = [&]() {
                if (cur.kind == Tok::Num) p.nums.push_back(cur.v);
                else if (cur.kind == Tok::Str) p.strs.push_back(cur.s);
                else if (cur.kind == Tok::Ident && (cur.s == "true" || cur.s == "false")) p.strs.push_back(cur.s);
                else throw Error("pbrt: bad parameter value for " + p.name);
                advance();
            };
            if (cur.kind == Tok::LBr) {
                advance();
                while (cur.kind != Tok::RBr) {
                    if (cur.kind == Tok::End) throw Error("pbrt: unterminated '['");
                    take();
                }
                advance();
            } else {
                take();
            }
            set.ps.push_back(std::move(p));
        }
        return set;
    }
};

struct GState {
    Mat4 ctm = Mat4::identity();
    std::string named_material; // current NamedMaterial
    std::optional<Material> material; // current anonymous Material
    bool has_area_light = false;
    Color area_light;
    bool reverse_orientation = false;
};

Mat4 mat_from_16(const std::vector<double> &v) {
    // pbrt lists the matrix column by column, which is also cgmath's storage order
    Mat4 m{};
    for (int i = 0; i < 16; i++) m.m[i] = (float)v[i];
    return m;
}

// pbrt_rs::ShapeInfo after to_trimesh(): object-space data + the CTM, orientation, material and emission at definition
struct RawShape {
    std::vector<float> P, N, uv;
    std::vector<uint32_t> idx;
    Mat4 matrix = Mat4::identity();
    bool reverse_orientation = false;
    Material bsdf;
    bool is_light = false;
    Color emission;
};
// PBRTSceneLoader::transform_mesh, scene_loader.rs:80-157: mat = matrix * m.matrix; points by transform_point, normals by
// transform_vector (negated first under ReverseOrientation); Mesh::new renormalises the normals (geometry.rs:143-162)
void emit_mesh(Scene &scene, const RawShape &rs, const Mat4 &matrix, bool use_shading_normal) {
    Mat4 mat = matrix * rs.matrix;
    auto mesh = std::make_shared<Mesh>();
    mesh->name = "noname"; // scene_loader.rs:137
    size_t nv = rs.P.size() / 3;
    for (size_t i = 0; i < nv; i++) {
        Vec3 q = mat.transform_point(Vec3{rs.P[3 * i], rs.P[3 * i + 1], rs.P[3 * i + 2]});
        mesh->vertices.insert(mesh->vertices.end(), {q.x, q.y, q.z});
    }
    mesh->indices = rs.idx;
    if (use_shading_normal && !rs.N.empty()) {
        size_t nb_wrong = 0;
        for (size_t i = 0; i < nv; i++) {
            Vec3 n{rs.N[3 * i], rs.N[3 * i + 1], rs.N[3 * i + 2]};
            if (rs.reverse_orientation) n = Vec3{-n.x, -n.y, -n.z};
            n = mat.transform_vector(n);
            float l = n.x * n.x + n.y * n.y + n.z * n.z;
            if (l == 0.0f) nb_wrong++;
            else if (l != 1.0f) {
                float s = std::sqrt(l);
                n = Vec3{n.x / s, n.y / s, n.z / s};
            }
            mesh->normals.insert(mesh->normals.end(), {n.x, n.y, n.z});
        }
        if (nb_wrong == nv) mesh->normals.clear(); // geometry.rs:159-162
    }
    mesh->uv = rs.uv;
    mesh->bsdf = rs.bsdf;
    if (rs.is_light) {
        mesh->is_light = true;
        mesh->emission = rs.emission;
    }
    if (!mesh->indices.empty()) scene.meshes.push_back(mesh); // geometry.rs:165-167
}

// Minimal PLY reader (ascii 1.0 / binary_little_endian 1.0): vertex x y z [nx ny nz] [u v | s t], faces as index lists,
// polygons fanned into triangles.  Stands in for pbrt_rs::ply::read_ply(..).to_trimesh() (crate not vendored: unpinned).
void read_ply(const std::string &filename, RawShape &out) {
    std::ifstream f(filename, std::ios::binary);
    if (!f) throw Error("ply: cannot open " + filename);
    std::string line;
    if (!std::getline(f, line) || line.substr(0, 3) != "ply") throw Error("ply: bad magic in " + filename);
    struct Prop {
        std::string name, type, count_type, item_type;
        bool list = false;
    };
    struct Elem {
        std::string name;
        size_t count = 0;
        std::vector<Prop> props;
    };
    std::vector<Elem> elems;
    bool binary = false, big_endian = false;
    while (std::getline(f, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        std::istringstream ls(line);
        std::string w;
        ls >> w;
        if (w == "format") {
            std::string fmt;
            ls >> fmt;
            if (fmt == "binary_little_endian") binary = true;
            else if (fmt == "binary_big_endian") binary = true, big_endian = true;
            else if (fmt != "ascii") throw Error("ply: unsupported format " + fmt);
        } else if (w == "element") {
            Elem e;
            ls >> e.name >> e.count;
            elems.push_back(e);
        } else if (w == "property") {
            if (elems.empty()) throw Error("ply: property before element");
            Prop p;
            ls >> p.type;
            if (p.type == "list") {
                p.list = true;
                ls >> p.count_type >> p.item_type >> p.name;
            } else ls >> p.name;
            elems.back().props.push_back(p);
        } else if (w == "end_header") break;
    }
    auto size_of = [](const std::string &t) -> size_t {
        if (t == "char" || t == "uchar" || t == "int8" || t == "uint8") return 1;
        if (t == "short" || t == "ushort" || t == "int16" || t == "uint16") return 2;
        if (t == "int" || t == "uint" || t == "float" || t == "int32" || t == "uint32" || t == "float32") return 4;
        if (t == "double" || t == "float64") return 8;
        throw Error("ply: unknown type " + t);
    };
    auto read_num = [&](const std::string &t) -> double {
        if (!binary) {
            double v;
            if (!(f >> v)) throw Error("ply: truncated ascii data");
            return v;
        }
        unsigned char b[8];
        size_t n = size_of(t);
        f.read(reinterpret_cast<char *>(b), (std::streamsize)n);
        if (!f) throw Error("ply: truncated binary data");
        if (big_endian) std::reverse(b, b + n);
        if (t == "float" || t == "float32") { float v; std::memcpy(&v, b, 4); return v; }
        if (t == "double" || t == "float64") { double v; std::memcpy(&v, b, 8); return v; }
        if (t == "char" || t == "int8") return (signed char)b[0];
        if (t == "uchar" || t == "uint8") return b[0];
        if (t == "short" || t == "int16") { int16_t v; std::memcpy(&v, b, 2); return v; }
        if (t == "ushort" || t == "uint16") { uint16_t v; std::memcpy(&v, b, 2); return v; }
        if (t == "int" || t == "int32") { int32_t v; std::memcpy(&v, b, 4); return v; }
        uint32_t v;
        std::memcpy(&v, b, 4);
        return v;
    };
    bool have_n = false, have_uv = false;
    for (auto &e : elems) {
        if (e.name == "vertex") {
            int ix = -1, iy = -1, iz = -1, inx = -1, iny = -1, inz = -1, iu = -1, iv = -1;
            for (size_t k = 0; k < e.props.size(); k++) {
                const std::string &n = e.props[k].name;
                if (e.props[k].list) throw Error("ply: list property on vertices");
                if (n == "x") ix = (int)k; else if (n == "y") iy = (int)k; else if (n == "z") iz = (int)k;
                else if (n == "nx") inx = (int)k; else if (n == "ny") iny = (int)k; else if (n == "nz") inz = (int)k;
                else if (n == "u" || n == "s") iu = (int)k; else if (n == "v" || n == "t") iv = (int)k;
            }
            if (ix < 0 || iy < 0 || iz < 0) throw Error("ply: vertex needs x y z");
            have_n = inx >= 0 && iny >= 0 && inz >= 0, have_uv = iu >= 0 && iv >= 0;
            std::vector<double> row(e.props.size());
            for (size_t i = 0; i < e.count; i++) {
                for (size_t k = 0; k < e.props.size(); k++) row[k] = read_num(e.props[k].type);
                out.P.insert(out.P.end(), {(float)row[ix], (float)row[iy], (float)row[iz]});
                if (have_n) out.N.insert(out.N.end(), {(float)row[inx], (float)row[iny], (float)row[inz]});
                if (have_uv) out.uv.insert(out.uv.end(), {(float)row[iu], (float)row[iv]});
            }
        } else if (e.name == "face") {
            for (size_t i = 0; i < e.count; i++)
                for (auto &p : e.props) {
                    if (!p.list) {
                        read_num(p.type);
                        continue;
                    }
                    size_t n = (size_t)read_num(p.count_type);
                    std::vector<uint32_t> poly(n);
                    for (size_t k = 0; k < n; k++) poly[k] = (uint32_t)read_num(p.item_type);
                    if (p.name != "vertex_indices" && p.name != "vertex_index") continue;
                    for (size_t k = 2; k < n; k++) out.idx.insert(out.idx.end(), {poly[0], poly[k - 1], poly[k]});
                }
        } else {
            for (size_t i = 0; i < e.count; i++)
                for (auto &p : e.props) {
                    if (!p.list) read_num(p.type);
                    else {
                        size_t n = (size_t)read_num(p.count_type);
                        for (size_t k = 0; k < n; k++) read_num(p.item_type);
                    }
                }
        }
    }
    size_t nv = out.P.size() / 3;
    for (uint32_t v : out.idx)
        if (v >= nv) throw Error("ply: face index out of range in " + filename);
}

// "texture <name>" "tex" -> BSDFColor::Bitmap (bsdf_texture_match_pbrt, bsdfs/mod.rs:224-236, applied to every colour parameter of
// bsdf_pbrt: Kd, Ks, Kr, Kt, eta, k); 0 when the parameter is not a texture
uint32_t texture_of(const ParamSet &ps, const char *name, const std::map<std::string, uint32_t> &texture_ids) {
    const Param *p = ps.find(name);
    if (!p || p->type != "texture") return 0;
    if (p->strs.empty() || !texture_ids.count(p->strs[0])) throw Error(std::string("pbrt: ") + name + " refers to an unknown texture"); // the reference warns / unwraps
    return texture_ids.at(p->strs[0]);
}
uint32_t kd_texture_of(const ParamSet &ps, const std::map<std::string, uint32_t> &texture_ids) { return texture_of(ps, "Kd", texture_ids); }
Material material_from_params(const std::string &type, const ParamSet &ps, const std::map<std::string, uint32_t> &texture_ids) {
    if (type == "matte") {
        // pbrt_rs::BSDF::Matte { kd } -> BSDFDiffuse (src/bsdfs/mod.rs:299-306); Kd default 0.5
        if (uint32_t t = kd_texture_of(ps, texture_ids)) {
            Material m = Material::diffuse(Color{0.0f, 0.0f, 0.0f});
            m.m.kd_texture = t;
            return m;
        }
        return Material::diffuse(ps.rgb("Kd", Color{0.5f, 0.5f, 0.5f}));
    }
    if (type == "phong") { // extension, see file header
        return Material::phong(ps.rgb("Kd", Color{0.5f, 0.5f, 0.5f}), ps.rgb("Ks", Color{0.5f, 0.5f, 0.5f}),
                               (float)ps.num("exponent", 30.0));
    }
    // distribution_pbrt (bsdfs/mod.rs:260-291): always GGX; isotropic only (distribution.rs:62 asserts alpha_u == alpha_v)
    auto ggx_alpha = [&](double def_roughness) {
        bool remap = ps.boolean("remaproughness", true);
        double r = ps.num("roughness", def_roughness);
        double u = ps.num("uroughness", r), v = ps.num("vroughness", r);
        if (u != v) throw Error("pbrt: anisotropic roughness panics in the reference (distribution.rs:62)");
        return remap_roughness((float)u, remap);
    };
    const Color black{0.0f, 0.0f, 0.0f};
    auto rgb_or_tex = [&](const char *name, Color def, uint32_t *tex) { // a textured slot keeps a black constant (it is never read)
        *tex = texture_of(ps, name, texture_ids);
        return *tex ? black : ps.rgb(name, def);
    };
    if (type == "mirror") { // bsdfs/mod.rs:349-357
        uint32_t t;
        Color kr = rgb_or_tex("Kr", Color{0.9f, 0.9f, 0.9f}, &t);
        Material m = Material::metal(kr, Color{1.0f, 1.0f, 1.0f}, Color{0.0f, 0.0f, 0.0f}, RL_MICROFACET_NONE, 0.0f);
        m.m.ks_texture = t;
        return m;
    }
    if (type == "metal") { // bsdfs/mod.rs:334-348; defaults: pbrt-v3's copper as RGB (pbrt_rs is not vendored: unpinned)
        uint32_t te, tk;
        Color eta = rgb_or_tex("eta", Color{0.2004f, 0.9240f, 1.1022f}, &te), k = rgb_or_tex("k", Color{3.9129f, 2.4528f, 2.1421f}, &tk);
        Material m = Material::metal(Color{1.0f, 1.0f, 1.0f}, eta, k, RL_MICROFACET_GGX, ggx_alpha(0.01));
        m.m.eta_texture = te, m.m.k_texture = tk;
        return m;
    }
    if (type == "glass") { // bsdfs/mod.rs:307-333: the distribution is ignored ("Pure glass instead"), .eta(eta, 1.0)
        uint32_t tr, tt;
        Color kr = rgb_or_tex("Kr", Color{1.0f, 1.0f, 1.0f}, &tr), kt = rgb_or_tex("Kt", Color{1.0f, 1.0f, 1.0f}, &tt);
        Material m = Material::glass(kr, kt, (float)ps.num("eta", ps.num("index", 1.5)), 1.0f);
        m.m.ks_texture = tr, m.m.kt_texture = tt;
        return m;
    }
    if (type == "substrate") { // bsdfs/mod.rs:358-374
        uint32_t t, ts;
        Color kd = rgb_or_tex("Kd", Color{0.5f, 0.5f, 0.5f}, &t), ks = rgb_or_tex("Ks", Color{0.5f, 0.5f, 0.5f}, &ts);
        Material m = Material::substrate(kd, ks, RL_MICROFACET_GGX, ggx_alpha(0.1));
        m.m.kd_texture = t, m.m.ks_texture = ts;
        return m;
    }
    throw Error("pbrt: material type \"" + type + "\" is not supported (matte, mirror, metal, glass, substrate, phong)");
}
} // namespace

Scene PBRTSceneLoader::load(const std::string &filename, bool use_shading_normal) const {
    size_t slash = filename.find_last_of('/');
    return load_string(read_file(filename), use_shading_normal, slash == std::string::npos ? "" : filename.substr(0, slash));
}

// `Include "file"` (one per line, as pbrt files write it): spliced in textually, relative to base_dir, before parsing
static std::string expand_includes(const std::string &text, const std::string &base_dir, int depth = 0) {
    if (depth > 16) throw Error("pbrt: Include nested deeper than 16 levels");
    if (text.find("Include") == std::string::npos) return text;
    std::istringstream in(text);
    std::string line, out;
    while (std::getline(in, line)) {
        std::string code = line.substr(0, line.find('#'));
        size_t a = code.find_first_not_of(" \t\r");
        if (a != std::string::npos && code.compare(a, 7, "Include") == 0) {
            size_t q0 = code.find('"', a + 7), q1 = q0 == std::string::npos ? q0 : code.find('"', q0 + 1);
            if (q1 == std::string::npos) throw Error("pbrt: Include needs a quoted file name");
            std::string fn = code.substr(q0 + 1, q1 - q0 - 1);
            if (!fn.empty() && fn[0] != '/' && !base_dir.empty()) fn = base_dir + "/" + fn;
            size_t slash = fn.find_last_of('/');
            out += expand_includes(read_file(fn), slash == std::string::npos ? base_dir : fn.substr(0, slash), depth + 1);
            out += "\n";
            out += code.substr(q1 + 1);
        } else out += line;
        out += "\n";
    }
    return out;
}

Scene PBRTSceneLoader::load_string(const std::string &text_in, bool use_shading_normal, const std::string &base_dir) const {
    const std::string text = expand_includes(text_in, base_dir);
    Parser ps(text);
    GState gs;
    std::vector<GState> stack;
    std::vector<Mat4> tstack;
    std::map<std::string, Mat4> named_cs;                   // CoordinateSystem "name"
    std::map<std::string, Material> materials;
    std::map<std::string, uint32_t> texture_ids;            // pbrt_rs::Scene::textures -> 1-based ids of scene.textures
    std::vector<RawShape> top_shapes;                       // pbrt_rs::Scene::shapes
    std::map<std::string, std::vector<RawShape>> objects;   // pbrt_rs::Scene::objects
    std::vector<std::pair<std::string, Mat4>> instances;    // pbrt_rs::Scene::instances {name, matrix}
    std::string current_object;
    Scene scene;
    uint32_t xres = 512, yres = 512;
    bool have_camera = false;
    float fov = 90.0f;
    Mat4 world_to_camera = Mat4::identity();

    while (ps.cur.kind != Tok::End) {
        if (ps.cur.kind != Tok::Ident) throw Error("pbrt: expected a directive");
        std::string d = ps.cur.s;
        ps.advance();
        if (d == "Transform") {
            gs.ctm = mat_from_16(ps.nums(16));
        } else if (d == "ConcatTransform") {
            gs.ctm = gs.ctm * mat_from_16(ps.nums(16));
        } else if (d == "Identity") {
            gs.ctm = Mat4::identity();
        } else if (d == "Translate") {
            auto v = ps.nums(3);
            gs.ctm = gs.ctm * Mat4::from_translation((float)v[0], (float)v[1], (float)v[2]);
        } else if (d == "Scale") {
            auto v = ps.nums(3);
            gs.ctm = gs.ctm * Mat4::from_nonuniform_scale((float)v[0], (float)v[1], (float)v[2]);
        } else if (d == "Rotate") {
            auto v = ps.nums(4);
            gs.ctm = gs.ctm * Mat4::rotate_deg((float)v[0], Vec3{(float)v[1], (float)v[2], (float)v[3]});
        } else if (d == "LookAt") {
            auto v = ps.nums(9);
            gs.ctm = gs.ctm * Mat4::look_at_pbrt(Vec3{(float)v[0], (float)v[1], (float)v[2]},
                                                 Vec3{(float)v[3], (float)v[4], (float)v[5]},
                                                 Vec3{(float)v[6], (float)v[7], (float)v[8]});
        } else if (d == "Film") {
            ps.expect_str();
            ParamSet p = ps.params();
            xres = (uint32_t)p.num("xresolution", 512);
            yres = (uint32_t)p.num("yresolution", 512);
        } else if (d == "Camera") {
            std::string type = ps.expect_str();
            ParamSet p = ps.params();
            if (type != "perspective") throw Error("pbrt: only Camera \"perspective\" is supported");
            fov = (float)p.num("fov", 90.0);
            world_to_camera = gs.ctm;
            if (auto inv = gs.ctm.invert()) named_cs["camera"] = *inv; // pbrt-v3: camera space -> world
            have_camera = true;
        } else if (d == "Integrator" || d == "Sampler" || d == "PixelFilter" || d == "Accelerator") {
            ps.expect_str();
            ps.params(); // rendering settings come from the CLI in rustlight, not from the file
        } else if (d == "WorldBegin") {
            gs = GState{};
        } else if (d == "WorldEnd") {
        } else if (d == "AttributeBegin") {
            stack.push_back(gs);
        } else if (d == "AttributeEnd") {
            if (stack.empty()) throw Error("pbrt: unbalanced AttributeEnd");
            gs = stack.back();
            stack.pop_back();
        } else if (d == "TransformBegin") {
            tstack.push_back(gs.ctm);
        } else if (d == "TransformEnd") {
            if (tstack.empty()) throw Error("pbrt: unbalanced TransformEnd");
            gs.ctm = tstack.back();
            tstack.pop_back();
        } else if (d == "CoordinateSystem") { // pbrt-v3: name the current transformation ("camera" is set by Camera)
            named_cs[ps.expect_str()] = gs.ctm;
        } else if (d == "CoordSysTransform") {
            const std::string name = ps.expect_str();
            auto it = named_cs.find(name);
            if (it == named_cs.end()) throw Error("pbrt: CoordSysTransform of unknown coordinate system " + name);
            gs.ctm = it->second;
        } else if (d == "ReverseOrientation") {
            gs.reverse_orientation = !gs.reverse_orientation;
        } else if (d == "MakeNamedMaterial") {
            std::string name = ps.expect_str();
            ParamSet p = ps.params();
            materials[name] = material_from_params(p.str("type", "matte"), p, texture_ids);
        } else if (d == "NamedMaterial") {
            gs.named_material = ps.expect_str();
            gs.material.reset();
        } else if (d == "Material") {
            std::string type = ps.expect_str();
            ParamSet p = ps.params();
            gs.material = material_from_params(type, p, texture_ids);
            gs.named_material.clear();
        } else if (d == "AreaLightSource") {
            std::string type = ps.expect_str();
            ParamSet p = ps.params();
            if (type != "diffuse" && type != "area") throw Error("pbrt: only AreaLightSource \"diffuse\"");
            gs.has_area_light = true;
            gs.area_light = p.rgb("L", Color{1.0f, 1.0f, 1.0f});
        } else if (d == "Shape") {
            std::string type = ps.expect_str();
            ParamSet p = ps.params();
            RawShape rs;
            if (type == "trianglemesh") {
                const Param *pi = p.find("indices"), *pp = p.find("P"), *pn = p.find("N"), *puv = p.find("uv");
                if (!puv) puv = p.find("st");
                if (!pi || !pp) throw Error("pbrt: trianglemesh needs indices and P");
                if (pi->nums.size() % 3 || pp->nums.size() % 3) throw Error("pbrt: trianglemesh sizes");
                for (double v : pp->nums) rs.P.push_back((float)v);
                for (double v : pi->nums) {
                    if (v < 0 || (size_t)v >= pp->nums.size() / 3) throw Error("pbrt: trianglemesh index out of range");
                    rs.idx.push_back((uint32_t)v);
                }
                if (pn) {
                    if (pn->nums.size() != pp->nums.size()) throw Error("pbrt: N size mismatch");
                    for (double v : pn->nums) rs.N.push_back((float)v);
                }
                if (puv) {
                    if (puv->nums.size() != 2 * (pp->nums.size() / 3)) throw Error("pbrt: uv size mismatch");
                    for (double v : puv->nums) rs.uv.push_back((float)v);
                }
            } else if (type == "plymesh") { // pbrt_rs::ply::read_ply(..).to_trimesh(), scene_loader.rs:88-92
                std::string fn = p.str("filename");
                if (fn.empty()) throw Error("pbrt: plymesh needs a filename");
                if (fn[0] != '/' && !base_dir.empty()) fn = base_dir + "/" + fn;
                read_ply(fn, rs);
            } else throw Error("pbrt: Shape \"" + type + "\" is outside the hot-path scope (trianglemesh, plymesh)");
            rs.matrix = gs.ctm;
            rs.reverse_orientation = gs.reverse_orientation;
            // scene_loader.rs:124-136: unknown / missing material -> diffuse 0.5
            if (gs.material) rs.bsdf = *gs.material;
            else if (!gs.named_material.empty() && materials.count(gs.named_material)) rs.bsdf = materials[gs.named_material];
            else rs.bsdf = Material::diffuse(Color{0.5f, 0.5f, 0.5f});
            rs.is_light = gs.has_area_light, rs.emission = gs.area_light; // scene_loader.rs:141-145
            if (!current_object.empty()) objects[current_object].push_back(std::move(rs));
            else top_shapes.push_back(std::move(rs));
        } else if (d == "ObjectBegin") { // pbrt_rs::Scene::objects (scene_loader.rs:183-203)
            if (!current_object.empty()) throw Error("pbrt: nested ObjectBegin");
            current_object = ps.expect_str();
            objects[current_object];
            stack.push_back(gs); // ObjectBegin implies AttributeBegin
            tstack.push_back(gs.ctm);
        } else if (d == "ObjectEnd") {
            if (current_object.empty() || stack.empty()) throw Error("pbrt: unbalanced ObjectEnd");
            current_object.clear();
            gs = stack.back();
            stack.pop_back();
            tstack.pop_back();
        } else if (d == "ObjectInstance") {
            std::string name = ps.expect_str();
            if (!objects.count(name)) throw Error("pbrt: ObjectInstance of unknown object " + name);
            instances.push_back({name, gs.ctm});
        } else if (d == "LightSource") { // scene_loader.rs:207-240 (pbrt_rs::Light::{Point, Distant}); positions in world space
            std::string type = ps.expect_str();
            ParamSet p = ps.params();
            Color scale = p.rgb("scale", Color{1.0f, 1.0f, 1.0f});
            auto pt = [&](const char *name, float dx, float dy, float dz) {
                auto q = p.find(name);
                Vec3 v = (q && q->nums.size() >= 3) ? Vec3{(float)q->nums[0], (float)q->nums[1], (float)q->nums[2]} : Vec3{dx, dy, dz};
                return gs.ctm.transform_point(v);
            };
            if (type == "point") {
                Color I = p.rgb("I", Color{1.0f, 1.0f, 1.0f});
                Vec3 from = pt("from", 0, 0, 0);
                scene.add_point_light(Color{I.r * scale.r, I.g * scale.g, I.b * scale.b}, from.x, from.y, from.z);
            } else if (type == "distant") {
                Color L = p.rgb("L", Color{1.0f, 1.0f, 1.0f});
                Vec3 from = pt("from", 0, 0, 0), to = pt("to", 0, 0, 1);
                scene.add_directional_light(Color{L.r * scale.r, L.g * scale.g, L.b * scale.b}, to.x - from.x, to.y - from.y, to.z - from.z);
            } else if (type == "infinite") { // scene_loader.rs:239-276: a constant RGB L (x scale), or a lat-long image ("mapname")
                if (scene.has_environment) throw Error("Multiple env map is NOT supported"); // scene_loader.rs:247-249
                if (p.find("mapname")) { // Spectrum::Mapname: EnvironmentLightColor::new_texture(Bitmap::read(file)); the reference asserts scale == 1
                    if (scale.r != 1.0f || scale.g != 1.0f || scale.b != 1.0f) throw Error("pbrt: LightSource \"infinite\" with a mapname needs scale 1 (scene_loader.rs:260-262)");
                    std::string fn = p.str("mapname");
                    if (fn.empty()) throw Error("pbrt: empty mapname");
                    if (fn[0] != '/' && !base_dir.empty()) fn = base_dir + "/" + fn;
                    scene.set_environment_texture(scene.add_texture(Texture::bitmap_file(fn)));
                } else {
                    Color L = p.rgb("L", Color{1.0f, 1.0f, 1.0f});
                    scene.set_environment(Color{L.r * scale.r, L.g * scale.g, L.b * scale.b});
                }
            } else throw Error("pbrt: LightSource \"" + type + "\" is outside the hot-path scope (point, distant, infinite)");
        } else if (d == "Texture") { // pbrt_rs::Texture {filename}: only image maps reach the BSDFs (bsdfs/mod.rs:219-241)
            std::string name = ps.expect_str(), ttype = ps.expect_str(), tclass = ps.expect_str();
            ParamSet p = ps.params();
            if (tclass != "imagemap") throw Error("pbrt: Texture class \"" + tclass + "\" is outside the hot-path scope (imagemap)");
            std::string fn = p.str("filename");
            if (fn.empty()) throw Error("pbrt: imagemap needs a filename");
            if (fn[0] != '/' && !base_dir.empty()) fn = base_dir + "/" + fn;
            texture_ids[name] = scene.add_texture(Texture::bitmap_file(fn));
            (void)ttype;
        } else if (d == "MakeNamedMedium" || d == "MediumInterface") {
            throw Error("pbrt: directive " + d + " is outside the hot-path scope");
        } else {
            throw Error("pbrt: unknown directive " + d);
        }
    }
    // scene_loader.rs:167-203: every top-level shape with the identity, then every instance's shapes with its matrix
    for (auto &rs : top_shapes) emit_mesh(scene, rs, Mat4::identity(), use_shading_normal);
    for (auto &in : instances)
        for (auto &rs : objects[in.first]) emit_mesh(scene, rs, in.second, use_shading_normal);
    if (!have_camera) throw Error("The camera is not set!"); // scene_loader.rs:295
    auto c2w = world_to_camera.invert();                     // scene_loader.rs:288
    if (!c2w) throw Error("pbrt: camera transform is singular");
    scene.camera = Camera::create(xres, yres, Fov::Y, fov, *c2w, false); // scene_loader.rs:291
    return scene;
}

// ------------------------------------------------------------------------------------------
// minimal JSON
// ------------------------------------------------------------------------------------------
namespace {
struct JVal {
    enum T { Null, Bool, Num, Str, Arr, Obj } t = Null;
    bool b = false;
    double n = 0;
    std::string s;
    std::vector<JVal> a;
    std::vector<std::pair<std::string, JVal>> o;
    const JVal *get(const std::string &k) const {
        for (auto &kv : o)
            if (kv.first == k) return &kv.second;
        return nullptr;
    }
};
struct JParser {
    const std::string &t;
    size_t p = 0;
    explicit JParser(const std::string &text) : t(text) {}
    void ws() {
        while (p < t.size() && std::isspace((unsigned char)t[p])) p++;
    }
    [[noreturn]] void fail(const std::string &m) { throw Error("json: " + m + " at byte " + std::to_string(p)); }
    JVal parse() {
        ws();
        if (p >= t.size()) fail("unexpected end");
        JVal v;
        char c = t[p];
        if (c == '{') {
            v.t = JVal::Obj;
            p++;
            ws();
            if (p < t.size() && t[p] == '}') { p++; return v; }
            for (;;) {
                ws();
                JVal k = parse();
                if (k.t != JVal::Str) fail("object key must be a string");
                ws();
                if (p >= t.size() || t[p] != ':') fail("expected ':'");
                p++;
                v.o.emplace_back(k.s, parse());
                ws();
                if (p < t.size() && t[p] == ',') { p++; continue; }
                if (p < t.size() && t[p] == '}') { p++; break; }
                fail("expected ',' or '}'");
            }
        } else if (c == '[') {
            v.t = JVal::Arr;
            p++;
            ws();
            if (p < t.size() && t[p] == ']') { p++; return v; }
            for (;;) {
                v.a.push_back(parse());
                ws();
                if (p < t.size() && t[p] == ',') { p++; continue; }
                if (p < t.size() && t[p] == ']') { p++; break; }
                fail("expected ',' or ']'");
            }
        } else if (c == '"') {
            v.t = JVal::Str;
            p++;
            while (p < t.size() && t[p] != '"') {
                if (t[p] == '\\' && p + 1 < t.size()) {
                    p++;
                    char e = t[p];
                    v.s.push_back(e == 'n' ? '\n' : e == 't' ? '\t' : e);
                } else v.s.push_back(t[p]);
                p++;
            }
            if (p >= t.size()) fail("unterminated string");
            p++;
        } else if (t.compare(p, 4, "true") == 0) { v.t = JVal::Bool; v.b = true; p += 4; }
        else if (t.compare(p, 5, "false") == 0) { v.t = JVal::Bool; v.b = false; p += 5; }
        else if (t.compare(p, 4, "null") == 0) { p += 4; }
        else {
            char *end = nullptr;
            v.n = std::strtod(t.c_str() + p, &end);
            if (end == t.c_str() + p) fail("bad value");
            v.t = JVal::Num;
            p = (size_t)(end - t.c_str());
        }
        return v;
    }
};
std::vector<float> jfloats(const JVal *v, const char *what, size_t expect = 0) {
    if (!v || v->t != JVal::Arr) throw Error(std::string("json: missing array ") + what);
    std::vector<float> r;
    for (auto &e : v->a) {
        if (e.t != JVal::Num) throw Error(std::string("json: non-number in ") + what);
        r.push_back((float)e.n);
    }
    if (expect && r.size() != expect) throw Error(std::string("json: wrong length for ") + what);
    return r;
}
Color jcolor(const JVal *v, const char *what, Color def) {
    if (!v) return def;
    auto f = jfloats(v, what, 3);
    return Color{f[0], f[1], f[2]};
}
Material jmaterial(const JVal &m) {
    const JVal *ty = m.get("type");
    std::string type = (ty && ty->t == JVal::Str) ? ty->s : "diffuse";
    if (type == "diffuse" || type == "matte") return Material::diffuse(jcolor(m.get("kd"), "kd", Color{0.5f, 0.5f, 0.5f}));
    if (type == "phong") {
        const JVal *e = m.get("exponent");
        return Material::phong(jcolor(m.get("kd"), "kd", Color{0.5f, 0.5f, 0.5f}),
                               jcolor(m.get("ks"), "ks", Color{0.5f, 0.5f, 0.5f}), e ? (float)e->n : 30.0f);
    }
    auto jnum = [&](const char *k, double def) {
        const JVal *v = m.get(k);
        return (v && v->t == JVal::Num) ? v->n : def;
    };
    auto jmicrofacet = [&]() -> uint32_t {
        const JVal *v = m.get("microfacet");
        std::string d = (v && v->t == JVal::Str) ? v->s : "ggx";
        if (d == "none") return RL_MICROFACET_NONE;
        if (d == "ggx") return RL_MICROFACET_GGX;
        if (d == "beckmann") return RL_MICROFACET_BECKMANN;
        throw Error("json: unknown microfacet \"" + d + "\"");
    };
    if (type == "mirror") return Material::metal(jcolor(m.get("ks"), "ks", Color{0.9f, 0.9f, 0.9f}), Color{1.0f, 1.0f, 1.0f}, Color{0.0f, 0.0f, 0.0f}, RL_MICROFACET_NONE, 0.0f);
    if (type == "metal")
        return Material::metal(jcolor(m.get("ks"), "ks", Color{1.0f, 1.0f, 1.0f}), jcolor(m.get("eta"), "eta", Color{0.2004f, 0.9240f, 1.1022f}),
                               jcolor(m.get("k"), "k", Color{3.9129f, 2.4528f, 2.1421f}), jmicrofacet(), (float)jnum("alpha", 0.1));
    if (type == "glass") { // "ior" = the relative index directly; else int_ior / ext_ior with BSDFGlass::default()'s bk7 / air (glass.rs:60-72)
        const bool rel = m.get("ior") != nullptr;
        return Material::glass(jcolor(m.get("ks"), "ks", Color{1.0f, 1.0f, 1.0f}), jcolor(m.get("kt"), "kt", Color{1.0f, 1.0f, 1.0f}),
                               (float)(rel ? jnum("ior", 1.5) : jnum("int_ior", 1.5046)), (float)(rel ? 1.0 : jnum("ext_ior", 1.000277)));
    }
    if (type == "substrate")
        return Material::substrate(jcolor(m.get("kd"), "kd", Color{0.5f, 0.5f, 0.5f}), jcolor(m.get("ks"), "ks", Color{0.5f, 0.5f, 0.5f}), jmicrofacet(), (float)jnum("alpha", 0.1));
    if (type == "blend") { // BSDFBlend (bsdfs/blend.rs): {"type": "blend", "a": {...}, "b": {...}, "weight": w}
        const JVal *a = m.get("a"), *b = m.get("b");
        if (!a || !b || a->t != JVal::Obj || b->t != JVal::Obj) throw Error("json: blend needs two material objects \"a\" and \"b\"");
        return Material::blend(jmaterial(*a), jmaterial(*b), (float)jnum("weight", 0.5));
    }
    throw Error("json: material type \"" + type + "\" is not supported (diffuse, phong, mirror, metal, glass, substrate, blend)");
}
} // namespace

Scene JSONSceneLoader::load(const std::string &filename, bool use_shading_normal) const {
    size_t slash = filename.find_last_of('/');
    return load_string(read_file(filename), use_shading_normal, slash == std::string::npos ? "" : filename.substr(0, slash));
}

Scene JSONSceneLoader::load_string(const std::string &text, bool use_shading_normal, const std::string &base_dir) const {
    JParser jp(text);
    JVal root = jp.parse();
    if (root.t != JVal::Obj) throw Error("json: root must be an object");
    Scene scene;
    const JVal *cam = root.get("camera");
    if (!cam || cam->t != JVal::Obj) throw Error("The camera is not set!");
    auto num = [](const JVal *v, double def) { return (v && v->t == JVal::Num) ? v->n : def; };
    uint32_t w = (uint32_t)num(cam->get("width"), 512), h = (uint32_t)num(cam->get("height"), 512);
    float fov = (float)num(cam->get("fov"), 90.0);
    const JVal *ax = cam->get("fov_axis");
    Fov axis = (ax && ax->t == JVal::Str && (ax->s == "x" || ax->s == "X")) ? Fov::X : Fov::Y;
    const JVal *fl = cam->get("flip");
    bool flip = fl && fl->t == JVal::Bool && fl->b;
    Mat4 to_world = Mat4::identity();
    if (cam->get("to_world")) {
        auto f = jfloats(cam->get("to_world"), "camera.to_world", 16);
        for (int i = 0; i < 16; i++) to_world.m[i] = f[i];
    }
    scene.camera = Camera::create(w, h, axis, fov, to_world, flip);

    // "textures": {name: {"type": "checkerboard", color0, color1, offset, scale} | {"type": "grid", ..., line_width} |
    //                      {"type": "bitmap", "filename": "x.pfm|x.ppm"} | {"type": "bitmap", "width", "height", "pixels": [r,g,b,...]}}
    std::map<std::string, uint32_t> texture_ids;
    if (const JVal *ts = root.get("textures")) {
        if (ts->t != JVal::Obj) throw Error("json: textures must be an object");
        for (auto &kv : ts->o) {
            const JVal &t = kv.second;
            const JVal *ty = t.get("type");
            std::string type = (ty && ty->t == JVal::Str) ? ty->s : "";
            auto f2 = [&](const char *k, float d0, float d1) {
                const JVal *v = t.get(k);
                if (!v) return std::vector<float>{d0, d1};
                return jfloats(v, k, 2);
            };
            if (type == "bitmap") {
                const JVal *fn = t.get("filename");
                if (fn && fn->t == JVal::Str) {
                    std::string path = fn->s;
                    if (path[0] != '/' && !base_dir.empty()) path = base_dir + "/" + path;
                    texture_ids[kv.first] = scene.add_texture(Texture::bitmap_file(path));
                } else {
                    const JVal *w = t.get("width"), *h = t.get("height");
                    if (!w || !h) throw Error("json: bitmap texture needs filename or width/height/pixels");
                    texture_ids[kv.first] = scene.add_texture(Texture::bitmap((uint32_t)w->n, (uint32_t)h->n, jfloats(t.get("pixels"), "pixels", 0)));
                }
            } else if (type == "checkerboard" || type == "grid") {
                Color c0 = jcolor(t.get("color0"), "color0", Color{0.4f, 0.4f, 0.4f}), c1 = jcolor(t.get("color1"), "color1", Color{0.2f, 0.2f, 0.2f});
                auto off = f2("offset", 0.0f, 0.0f), sc = f2("scale", 1.0f, 1.0f);
                const JVal *lw = t.get("line_width");
                texture_ids[kv.first] = scene.add_texture(type == "grid" ? Texture::grid(c0, c1, lw ? (float)lw->n : 0.01f, off[0], off[1], sc[0], sc[1])
                                                                          : Texture::checkerboard(c0, c1, off[0], off[1], sc[0], sc[1]));
            } else throw Error("json: unknown texture type \"" + type + "\"");
        }
    }
    auto with_texture = [&](const JVal &m, Material mat) {
        if (const JVal *kt = m.get("kd_texture")) {
            if (kt->t != JVal::Str || !texture_ids.count(kt->s)) throw Error("json: unknown kd_texture");
            if (mat.m.kind != RL_BSDF_DIFFUSE && mat.m.kind != RL_BSDF_PHONG && mat.m.kind != RL_BSDF_SUBSTRATE) throw Error("json: kd_texture on a material without a diffuse slot");
            mat.m.kd_texture = texture_ids[kt->s];
        }
        struct Slot {
            const char *key;
            uint32_t rl_material::*field;
            bool ok;
        } slots[] = {{"ks_texture", &rl_material::ks_texture, mat.m.kind != RL_BSDF_DIFFUSE},
                     {"kt_texture", &rl_material::kt_texture, mat.m.kind == RL_BSDF_GLASS},
                     {"eta_texture", &rl_material::eta_texture, mat.m.kind == RL_BSDF_METAL},
                     {"k_texture", &rl_material::k_texture, mat.m.kind == RL_BSDF_METAL}};
        for (const Slot &sl : slots)
            if (const JVal *kt = m.get(sl.key)) {
                if (kt->t != JVal::Str || !texture_ids.count(kt->s)) throw Error(std::string("json: unknown ") + sl.key);
                if (!sl.ok) throw Error(std::string("json: ") + sl.key + " on a material without that slot");
                mat.m.*(sl.field) = texture_ids[kt->s];
            }
        return mat;
    };
    std::map<std::string, Material> materials;
    if (const JVal *ms = root.get("materials")) {
        if (ms->t != JVal::Obj) throw Error("json: materials must be an object");
        for (auto &kv : ms->o) materials[kv.first] = with_texture(kv.second, jmaterial(kv.second));
    }
    if (const JVal *env = root.get("environment")) { // "environment": [r, g, b]: constant EnvironmentLight; {"texture": name}: a lat-long bitmap of "textures"
        if (env->t == JVal::Obj) {
            const JVal *tn = env->get("texture");
            if (!tn || tn->t != JVal::Str || !texture_ids.count(tn->s)) throw Error("json: environment.texture must name a bitmap texture");
            scene.set_environment_texture(texture_ids[tn->s]);
        } else {
            Color c = jcolor(env, "environment", Color{0.0f, 0.0f, 0.0f});
            scene.set_environment(c);
        }
    }
    if (const JVal *ls = root.get("lights")) { // [{"type": "point", "intensity": [r,g,b], "position": [x,y,z]}, {"type": "directional", "intensity", "direction"}]
        if (ls->t != JVal::Arr) throw Error("json: lights must be an array");
        for (auto &l : ls->a) {
            const JVal *ty = l.get("type");
            std::string type = (ty && ty->t == JVal::Str) ? ty->s : "";
            Color I = jcolor(l.get("intensity"), "intensity", Color{1.0f, 1.0f, 1.0f});
            if (type == "point") {
                auto p = jfloats(l.get("position"), "position", 3);
                scene.add_point_light(I, p[0], p[1], p[2]);
            } else if (type == "directional") {
                auto d = jfloats(l.get("direction"), "direction", 3);
                scene.add_directional_light(I, d[0], d[1], d[2]);
            } else throw Error("json: unknown light type \"" + type + "\"");
        }
    }
    const JVal *meshes = root.get("meshes");
    if (!meshes || meshes->t != JVal::Arr) throw Error("json: missing meshes array");
    for (auto &jm : meshes->a) {
        auto mesh = std::make_shared<Mesh>();
        const JVal *nm = jm.get("name");
        mesh->name = (nm && nm->t == JVal::Str) ? nm->s : "noname";
        mesh->vertices = jfloats(jm.get("P"), "mesh.P");
        if (mesh->vertices.size() % 3) throw Error("json: mesh.P length must be a multiple of 3");
        size_t nv = mesh->vertices.size() / 3;
        const JVal *idx = jm.get("indices");
        if (!idx || idx->t != JVal::Arr || idx->a.size() % 3) throw Error("json: mesh.indices");
        for (auto &e : idx->a) {
            if (e.t != JVal::Num || e.n < 0 || (size_t)e.n >= nv) throw Error("json: mesh index out of range");
            mesh->indices.push_back((uint32_t)e.n);
        }
        if (use_shading_normal && jm.get("N")) {
            mesh->normals = jfloats(jm.get("N"), "mesh.N", 3 * nv);
            size_t nb_wrong = 0;
            for (size_t i = 0; i < nv; i++) { // Mesh::new, geometry.rs:143-153
                float *n = &mesh->normals[3 * i];
                float l = n[0] * n[0] + n[1] * n[1] + n[2] * n[2];
                if (l == 0.0f) nb_wrong++;
                else if (l != 1.0f) {
                    float s = std::sqrt(l);
                    n[0] /= s, n[1] /= s, n[2] /= s;
                }
            }
            if (nb_wrong == nv) mesh->normals.clear();
        }
        if (jm.get("uv")) mesh->uv = jfloats(jm.get("uv"), "mesh.uv", 2 * nv);
        const JVal *mat = jm.get("material");
        if (mat && mat->t == JVal::Str) {
            if (!materials.count(mat->s)) throw Error("json: unknown material " + mat->s);
            mesh->bsdf = materials[mat->s];
        } else if (mat && mat->t == JVal::Obj) mesh->bsdf = with_texture(*mat, jmaterial(*mat));
        else mesh->bsdf = Material::diffuse(Color{0.5f, 0.5f, 0.5f});
        if (jm.get("emission")) {
            mesh->is_light = true;
            mesh->emission = jcolor(jm.get("emission"), "mesh.emission", Color{});
        }
        // EmissionType::HSV / Texture (geometry.rs:99-104): "emission_hsv": scale | "emission_texture": {"texture": name, "scale": s}
        if (const JVal *h = jm.get("emission_hsv")) {
            const float sc = jfloats(h, "mesh.emission_hsv", 1)[0];
            mesh->is_light = true, mesh->emission_kind = RL_EMISSION_HSV, mesh->emission = Color{sc, sc, sc};
        } else if (const JVal *t = jm.get("emission_texture")) {
            const JVal *tn = t->get("texture"), *ts = t->get("scale");
            if (!tn || tn->t != JVal::Str || !texture_ids.count(tn->s)) throw Error("json: mesh.emission_texture needs {\"texture\": <name of a texture>, \"scale\": s}");
            const uint32_t id = texture_ids[tn->s];
            if (scene.textures[id - 1].t.kind != RL_TEX_BITMAP) throw Error("json: mesh.emission_texture must name a bitmap texture");
            const float sc = ts ? jfloats(ts, "mesh.emission_texture.scale", 1)[0] : 1.0f;
            mesh->is_light = true, mesh->emission_kind = RL_EMISSION_TEXTURE, mesh->emission = Color{sc, sc, sc}, mesh->emission_texture = id;
        }
        if (!mesh->indices.empty()) scene.meshes.push_back(mesh);
    }
    return scene;
}

static void put_floats(std::ostringstream &o, const float *v, size_t n) {
    o << "[";
    char buf[32];
    for (size_t i = 0; i < n; i++) {
        std::snprintf(buf, sizeof(buf), "%.9g", (double)v[i]);
        o << (i ? ", " : "") << buf;
    }
    o << "]";
}

std::string scene_to_json(const Scene &scene) {
    std::ostringstream o;
    const Camera &c = scene.camera;
    o << "{\n  \"camera\": {\"width\": " << c.img_x << ", \"height\": " << c.img_y << ", \"fov\": ";
    char buf[32];
    std::snprintf(buf, sizeof(buf), "%.9g", (double)c.fov_deg);
    o << buf << ", \"fov_axis\": \"" << (c.fov_axis == Fov::X ? "x" : "y") << "\", \"flip\": "
      << (c.flip ? "true" : "false") << ",\n             \"to_world\": ";
    put_floats(o, c.to_world.m, 16);
    o << "},\n";
    if (scene.has_environment && scene.environment_texture) o << "  \"environment\": {\"texture\": \"tex" << scene.environment_texture << "\"},\n";
    else if (scene.has_environment) {
        float e[3] = {scene.environment.r, scene.environment.g, scene.environment.b};
        o << "  \"environment\": ";
        put_floats(o, e, 3);
        o << ",\n";
    }
    if (!scene.lights.empty()) {
        o << "  \"lights\": [";
        for (size_t i = 0; i < scene.lights.size(); i++) {
            const rl_light_desc &l = scene.lights[i];
            o << (i ? ", " : "") << "{\"type\": \"" << (l.kind == RL_LIGHT_POINT ? "point" : "directional") << "\", \"intensity\": ";
            put_floats(o, l.intensity, 3);
            o << ", \"" << (l.kind == RL_LIGHT_POINT ? "position" : "direction") << "\": ";
            put_floats(o, l.v, 3);
            o << "}";
        }
        o << "],\n";
    }
    if (!scene.textures.empty()) {
        o << "  \"textures\": {";
        for (size_t i = 0; i < scene.textures.size(); i++) {
            const Texture &t = scene.textures[i];
            o << (i ? ", " : "") << "\"tex" << i + 1 << "\": {\"type\": \"" << (t.t.kind == RL_TEX_BITMAP ? "bitmap" : (t.t.kind == RL_TEX_GRID ? "grid" : "checkerboard")) << "\"";
            if (t.t.kind == RL_TEX_BITMAP) {
                o << ", \"width\": " << t.t.width << ", \"height\": " << t.t.height << ", \"pixels\": ";
                put_floats(o, t.pixels.data(), t.pixels.size());
            } else {
                o << ", \"color0\": ";
                put_floats(o, t.t.color0, 3);
                o << ", \"color1\": ";
                put_floats(o, t.t.color1, 3);
                o << ", \"offset\": ";
                put_floats(o, t.t.offset, 2);
                o << ", \"scale\": ";
                put_floats(o, t.t.scale, 2);
                if (t.t.kind == RL_TEX_GRID) {
                    char b1[32];
                    std::snprintf(b1, sizeof(b1), "%.9g", (double)t.t.line_width);
                    o << ", \"line_width\": " << b1;
                }
            }
            o << "}";
        }
        o << "},\n";
    }
    o << "  \"meshes\": [\n";
    for (size_t i = 0; i < scene.meshes.size(); i++) {
        const Mesh &m = *scene.meshes[i];
        static const char *kinds[] = {"diffuse", "phong", "metal", "glass", "substrate", "blend"};
        static const char *mfs[] = {"none", "ggx", "beckmann"};
        auto put = [&](const char *key, const float *v, size_t n) {
            o << ", \"" << key << "\": ";
            if (n == 1) { // scalars are bare numbers
                char b1[32];
                std::snprintf(b1, sizeof(b1), "%.9g", (double)v[0]);
                o << b1;
            } else put_floats(o, v, n);
        };
        auto put_material = [&](const rl_material &mt) {
            o << "{\"type\": \"" << kinds[mt.kind <= RL_BSDF_BLEND ? mt.kind : 0] << "\"";
            if (mt.kind == RL_BSDF_DIFFUSE || mt.kind == RL_BSDF_PHONG || mt.kind == RL_BSDF_SUBSTRATE) put("kd", mt.kd, 3);
            if (mt.kind != RL_BSDF_DIFFUSE && mt.kind != RL_BSDF_BLEND) put("ks", mt.ks, 3);
            if (mt.kind == RL_BSDF_PHONG) put("exponent", &mt.exponent, 1);
            if (mt.kind == RL_BSDF_METAL) put("eta", mt.eta, 3), put("k", mt.k, 3);
            if (mt.kind == RL_BSDF_GLASS) put("kt", mt.kt, 3), put("ior", &mt.ior, 1);
            if (mt.kd_texture) o << ", \"kd_texture\": \"tex" << mt.kd_texture << "\"";
            if (mt.ks_texture) o << ", \"ks_texture\": \"tex" << mt.ks_texture << "\"";
            if (mt.kt_texture) o << ", \"kt_texture\": \"tex" << mt.kt_texture << "\"";
            if (mt.eta_texture) o << ", \"eta_texture\": \"tex" << mt.eta_texture << "\"";
            if (mt.k_texture) o << ", \"k_texture\": \"tex" << mt.k_texture << "\"";
            if (mt.kind == RL_BSDF_METAL || mt.kind == RL_BSDF_SUBSTRATE) {
                o << ", \"microfacet\": \"" << mfs[mt.microfacet <= RL_MICROFACET_BECKMANN ? mt.microfacet : 0] << "\"";
                put("alpha", &mt.alpha, 1);
            }
        };
        o << "    {\"name\": \"" << m.name << "\", \"material\": ";
        put_material(m.bsdf.m);
        if (m.bsdf.m.kind == RL_BSDF_BLEND && m.bsdf.subs.size() == 2) { // BSDFBlend { bsdf1: a, bsdf2: b, weight }
            o << ", \"a\": ";
            put_material(m.bsdf.subs[0]);
            o << "}, \"b\": ";
            put_material(m.bsdf.subs[1]);
            o << "}";
            put("weight", &m.bsdf.m.blend_weight, 1);
        }
        o << "}";
        if (m.is_light && m.emission_kind == RL_EMISSION_HSV) {
            o << ", \"emission_hsv\": ";
            put_floats(o, &m.emission.r, 1);
        } else if (m.is_light && m.emission_kind == RL_EMISSION_TEXTURE) {
            o << ", \"emission_texture\": {\"texture\": \"tex" << m.emission_texture << "\", \"scale\": ";
            put_floats(o, &m.emission.r, 1);
            o << "}";
        } else if (m.is_light) {
            float e[3] = {m.emission.r, m.emission.g, m.emission.b};
            o << ", \"emission\": ";
            put_floats(o, e, 3);
        }
        o << ",\n     \"indices\": [";
        for (size_t k = 0; k < m.indices.size(); k++) o << (k ? ", " : "") << m.indices[k];
        o << "],\n     \"P\": ";
        put_floats(o, m.vertices.data(), m.vertices.size());
        if (!m.normals.empty()) {
            o << ",\n     \"N\": ";
            put_floats(o, m.normals.data(), m.normals.size());
        }
        if (!m.uv.empty()) {
            o << ",\n     \"uv\": ";
            put_floats(o, m.uv.data(), m.uv.size());
        }
        o << "}" << (i + 1 < scene.meshes.size() ? "," : "") << "\n";
    }
    o << "  ]\n}\n";
    return o.str();
}

// ------------------------------------------------------------------------------------------
// MTSSceneLoader, src/scene_loader.rs:318-795 + bsdf_mts, src/bsdfs/mod.rs:395-612 (Mitsuba 0.x XML subset)
//
// The reference parses the file with the mitsuba_rs crate (git dependency, not vendored: its defaults are restated from the
// Mitsuba 0.5/0.6 documentation and are UNPINNED) and maps the result as follows, which is what this loader reproduces:
//   sensor "perspective": film width / height, fov + fovAxis x | y, toWorld -> Camera::new(size, fov, mat, flip = true)   (:327-338)
//   shapes "rectangle" (two triangles, normals, uv; :538-594), "sphere" (32 x 32 lat-long tessellation; :596-665), "ply", "obj", "serialized";
//     toWorld applied to points (transform_point) and normals (transform_vector, renormalised) (:341-376); bsdf (inline or <ref>),
//     default BSDFDiffuse 0.8; <emitter type="area"> radiance -> EmissionType::Color; faceNormals discards normals
//   bsdfs: twosided (ignored wrapper), diffuse, phong (weight_specular from the average luminances), dielectric -> BSDFGlass.eta(int, ext),
//     plastic / roughplastic -> BSDFSubstrate, conductor / roughconductor -> BSDFMetal (eta, k divided by extEta), distribution
//     "beckmann" | "ggx" with an isotropic alpha; anything else -> BSDFDiffuse 0.8                                           (mod.rs:499-612)
//   colours: <rgb>/<spectrum> constants, <texture type="bitmap" | "checkerboard" | "gridtexture">                           (mod.rs:395-452)
//   emitters "point" -> PointEmitter (:680-697).  The reference's "PointNormal" emitter cannot be sampled by `path` / `direct`
//     (PointNormalEmitter::direct_sample is todo!(), emitter.rs:262-264): rejected here instead of panicking at the first light sample.
//   media -> outside the hot path.  Shapes "serialized" (:499-538): read_serialized below (deflate through zlib).
// ------------------------------------------------------------------------------------------
namespace {
struct XNode {
    std::string tag;
    std::map<std::string, std::string> attr;
    std::vector<XNode> kids;
    const XNode *child(const std::string &tag_, const std::string &name) const {
        for (auto &k : kids)
            if (k.tag == tag_ && k.get("name") == name) return &k;
        return nullptr;
    }
    const XNode *named(const std::string &name) const {
        for (auto &k : kids)
            if (k.get("name") == name) return &k;
        return nullptr;
    }
    std::string get(const std::string &k, const std::string &def = "") const {
        auto it = attr.find(k);
        return it == attr.end() ? def : it->second;
    }
};
struct XParser {
    const std::string &t;
    size_t i = 0;
    std::map<std::string, std::string> defaults; // <default name= value=> for $name substitution
    explicit XParser(const std::string &text) : t(text) {}
    [[noreturn]] void fail(const std::string &m) const { throw Error("xml: " + m + " at offset " + std::to_string(i)); }
    void skip_misc() {
        for (;;) {
            while (i < t.size() && std::isspace((unsigned char)t[i])) i++;
            if (t.compare(i, 4, "<!--") == 0) {
                size_t e = t.find("-->", i);
                if (e == std::string::npos) fail("unterminated comment");
                i = e + 3;
            } else if (t.compare(i, 2, "<?") == 0) {
                size_t e = t.find("?>", i);
                if (e == std::string::npos) fail("unterminated declaration");
                i = e + 2;
            } else break;
        }
    }
    std::string subst(std::string v) const {
        for (auto &d : defaults) {
            const std::string key = "$" + d.first;
            for (size_t p = v.find(key); p != std::string::npos; p = v.find(key, p + d.second.size())) v.replace(p, key.size(), d.second);
        }
        return v;
    }
    XNode element() {
        skip_misc();
        if (i >= t.size() || t[i] != '<') fail("expected an element");
        i++;
        XNode n;
        while (i < t.size() && (std::isalnum((unsigned char)t[i]) || t[i] == '_' || t[i] == ':')) n.tag += t[i++];
        if (n.tag.empty()) fail("empty tag name");
        for (;;) {
            while (i < t.size() && std::isspace((unsigned char)t[i])) i++;
            if (i >= t.size()) fail("unterminated tag");
            if (t[i] == '/') {
                if (t.compare(i, 2, "/>") != 0) fail("bad tag end");
                i += 2;
                finish(n);
                return n;
            }
            if (t[i] == '>') {
                i++;
                break;
            }
            std::string key;
            while (i < t.size() && t[i] != '=' && !std::isspace((unsigned char)t[i])) key += t[i++];
            while (i < t.size() && std::isspace((unsigned char)t[i])) i++;
            if (i >= t.size() || t[i] != '=') fail("attribute without a value");
            i++;
            while (i < t.size() && std::isspace((unsigned char)t[i])) i++;
            const char q = i < t.size() ? t[i] : 0;
            if (q != '"' && q != '\'') fail("attribute value must be quoted");
            size_t e = t.find(q, i + 1);
            if (e == std::string::npos) fail("unterminated attribute value");
            n.attr[key] = subst(t.substr(i + 1, e - i - 1));
            i = e + 1;
        }
        for (;;) {
            skip_misc();
            if (i >= t.size()) fail("unterminated element <" + n.tag + ">");
            if (t.compare(i, 2, "</") == 0) {
                size_t e = t.find('>', i);
                if (e == std::string::npos) fail("unterminated end tag");
                i = e + 1;
                finish(n);
                return n;
            }
            if (t[i] != '<') { // text content: not used by the format
                i++;
                continue;
            }
            n.kids.push_back(element());
        }
    }
    void finish(const XNode &n) {
        if (n.tag == "default") defaults[n.get("name")] = n.get("value");
    }
};
std::vector<float> xfloats(const std::string &v) {
    std::vector<float> out;
    std::string cur;
    auto flush = [&]() {
        if (!cur.empty()) out.push_back(std::strtof(cur.c_str(), nullptr)), cur.clear();
    };
    for (char c : v) {
        if (std::isspace((unsigned char)c) || c == ',') flush();
        else cur += c;
    }
    flush();
    return out;
}
float xfloat(const XNode &n, const std::string &name, float def) {
    const XNode *c = n.child("float", name);
    if (!c) c = n.child("integer", name);
    return c ? std::strtof(c->get("value").c_str(), nullptr) : def;
}
bool xbool(const XNode &n, const std::string &name, bool def) {
    const XNode *c = n.child("boolean", name);
    return c ? c->get("value") == "true" : def;
}
std::string xstring(const XNode &n, const std::string &name, const std::string &def) {
    const XNode *c = n.child("string", name);
    return c ? c->get("value") : def;
}
// Spectrum::as_rgb(): <rgb value="r, g, b"> | <rgb value="v"> | <spectrum value="v"> (a constant; wavelength lists are not converted)
Color xrgb_value(const XNode &c) {
    if (c.tag != "rgb" && c.tag != "spectrum" && c.tag != "srgb") throw Error("xml: <" + c.tag + " name=\"" + c.get("name") + "\"> is not a colour");
    std::vector<float> v = xfloats(c.get("value"));
    if (c.get("value").find(':') != std::string::npos) throw Error("xml: spectra given as wavelength:value lists are not supported");
    if (v.size() == 1) return Color{v[0], v[0], v[0]};
    if (v.size() == 3) return Color{v[0], v[1], v[2]};
    throw Error("xml: a colour needs 1 or 3 values");
}
Color xrgb(const XNode &n, const std::string &name, Color def) {
    const XNode *c = n.named(name);
    return c ? xrgb_value(*c) : def;
}
// to_world.as_matrix(): the children of <transform> in order, each applied AFTER the previous ones (M = op * M)
Mat4 xtransform(const XNode *tr) {
    Mat4 m = Mat4::identity();
    if (!tr) return m;
    for (auto &op : tr->kids) {
        Mat4 o = Mat4::identity();
        auto f = [&](const char *k, float d) { return op.attr.count(k) ? std::strtof(op.get(k).c_str(), nullptr) : d; };
        if (op.tag == "matrix") { // row-major in the file
            std::vector<float> v = xfloats(op.get("value"));
            if (v.size() != 16) throw Error("xml: <matrix> needs 16 values");
            for (int r = 0; r < 4; r++)
                for (int c = 0; c < 4; c++) o.at(c, r) = v[4 * r + c];
        } else if (op.tag == "translate") o = Mat4::from_translation(f("x", 0), f("y", 0), f("z", 0));
        else if (op.tag == "scale") {
            const float u = f("value", 1.0f);
            o = Mat4::from_nonuniform_scale(f("x", u), f("y", u), f("z", u));
        } else if (op.tag == "rotate") o = Mat4::rotate_deg(f("angle", 0), Vec3{f("x", 0), f("y", 0), f("z", 0)});
        else if (op.tag == "lookat" || op.tag == "lookAt") { // camera-to-world: columns left, up, dir, origin (Mitsuba's Transform::lookAt)
            std::vector<float> e = xfloats(op.get("origin")), a = xfloats(op.get("target")), u = xfloats(op.get("up", "0, 1, 0"));
            if (e.size() != 3 || a.size() != 3 || u.size() != 3) throw Error("xml: <lookat> needs origin, target (and up) with 3 values");
            auto norm = [](Vec3 v) {
                float l = std::sqrt(v.x * v.x + v.y * v.y + v.z * v.z);
                return Vec3{v.x / l, v.y / l, v.z / l};
            };
            auto cross = [](Vec3 p, Vec3 q) { return Vec3{p.y * q.z - p.z * q.y, p.z * q.x - p.x * q.z, p.x * q.y - p.y * q.x}; };
            Vec3 dir = norm(Vec3{a[0] - e[0], a[1] - e[1], a[2] - e[2]});
            Vec3 left = norm(cross(Vec3{u[0], u[1], u[2]}, dir));
            Vec3 up = cross(dir, left);
            const float cols[4][4] = {{left.x, left.y, left.z, 0}, {up.x, up.y, up.z, 0}, {dir.x, dir.y, dir.z, 0}, {e[0], e[1], e[2], 1}};
            for (int c = 0; c < 4; c++)
                for (int r = 0; r < 4; r++) o.at(c, r) = cols[c][r];
        } else throw Error("xml: transform operation <" + op.tag + "> is not supported");
        m = o * m;
    }
    return m;
}
struct MtsCtx {
    Scene *scene;
    std::string base_dir;
    std::map<std::string, const XNode *> ids; // id -> bsdf / texture node
};
// bsdf_texture_match_mts (mod.rs:395-452): constant -> (colour, 0); texture -> (black, 1-based texture id)
std::pair<Color, uint32_t> mts_color(MtsCtx &cx, const XNode &owner, const std::string &name, Color def) {
    const XNode *c = owner.named(name);
    if (!c) return {def, 0u};
    if (c->tag == "ref") {
        auto it = cx.ids.find(c->get("id"));
        if (it == cx.ids.end()) throw Error("xml: <ref id=\"" + c->get("id") + "\"> refers to nothing");
        c = it->second;
    }
    if (c->tag != "texture") return {xrgb_value(*c), 0u};
    const std::string type = c->get("type");
    if (type == "bitmap") {
        std::string fn = xstring(*c, "filename", "");
        if (fn.empty()) throw Error("xml: bitmap texture without a filename");
        if (fn[0] != '/' && !cx.base_dir.empty()) fn = cx.base_dir + "/" + fn;
        Texture t = Texture::bitmap_file(fn);
        const float gamma = xfloat(*c, "gamma", 1.0f);
        if (gamma != 1.0f) // img.gamma(1.0 / gamma): per channel powf (structure.rs:416-422)
            for (float &v : t.pixels) v = std::pow(v, 1.0f / gamma);
        t.t.pixels = t.pixels.data();
        return {Color{0, 0, 0}, cx.scene->add_texture(std::move(t))};
    }
    if (type == "checkerboard" || type == "gridtexture") {
        Color c0 = xrgb(*c, "color0", type == "checkerboard" ? Color{0.4f, 0.4f, 0.4f} : Color{0.2f, 0.2f, 0.2f});
        Color c1 = xrgb(*c, "color1", type == "checkerboard" ? Color{0.2f, 0.2f, 0.2f} : Color{0.4f, 0.4f, 0.4f});
        const float ox = xfloat(*c, "uoffset", 0.0f), oy = xfloat(*c, "voffset", 0.0f), sx = xfloat(*c, "uscale", 1.0f), sy = xfloat(*c, "vscale", 1.0f);
        if (type == "checkerboard") return {Color{0, 0, 0}, cx.scene->add_texture(Texture::checkerboard(c0, c1, ox, oy, sx, sy))};
        return {Color{0, 0, 0}, cx.scene->add_texture(Texture::grid(c0, c1, xfloat(*c, "lineWidth", 0.01f), ox, oy, sx, sy))};
    }
    throw Error("Mitsuba texture type not supported: " + type); // mod.rs:450
}
uint32_t mts_distribution(const XNode &b, bool rough, float *alpha) { // distribution_mts, mod.rs:462-497
    *alpha = 0.0f;
    if (!rough) return RL_MICROFACET_NONE;
    if (b.child("float", "alphaU") || b.child("float", "alphaV")) {
        const float au = xfloat(b, "alphaU", 0.1f), av = xfloat(b, "alphaV", 0.1f);
        if (au != av) throw Error("xml: anisotropic roughness is not supported (the reference asserts alpha_u == alpha_v, mod.rs:476)");
        *alpha = au;
    } else *alpha = xfloat(b, "alpha", 0.1f);
    const std::string d = xstring(b, "distribution", "beckmann");
    return d == "ggx" ? RL_MICROFACET_GGX : RL_MICROFACET_BECKMANN; // unknown names: "Unsupported microfacet type" -> Beckmann (mod.rs:484-487)
}
Material mts_bsdf(MtsCtx &cx, const XNode *b) { // bsdf_mts, mod.rs:499-612
    const Material fallback = Material::diffuse(Color{0.8f, 0.8f, 0.8f});
    if (!b) return fallback;
    if (b->tag == "ref") {
        auto it = cx.ids.find(b->get("id"));
        if (it == cx.ids.end()) throw Error("xml: <ref id=\"" + b->get("id") + "\"> refers to nothing");
        b = it->second;
    }
    const std::string type = b->get("type");
    auto textured = [](Material m, uint32_t kd_t, uint32_t ks_t) {
        m.m.kd_texture = kd_t, m.m.ks_texture = ks_t;
        return m;
    };
    if (type == "twosided") { // "Rustlight automatically apply twosided"
        for (auto &k : b->kids)
            if (k.tag == "bsdf" || k.tag == "ref") return mts_bsdf(cx, &k);
        return fallback;
    }
    if (type == "diffuse") {
        auto r = mts_color(cx, *b, "reflectance", Color{0.5f, 0.5f, 0.5f});
        return textured(Material::diffuse(r.first), r.second, 0u);
    }
    if (type == "phong") {
        auto kd = mts_color(cx, *b, "diffuseReflectance", Color{0.5f, 0.5f, 0.5f}), ks = mts_color(cx, *b, "specularReflectance", Color{0.2f, 0.2f, 0.2f});
        if (kd.second || ks.second) throw Error("xml: textured phong reflectances are not supported (weight_specular needs BSDFColor::avg)");
        return Material::phong(kd.first, ks.first, xfloat(*b, "exponent", 30.0f));
    }
    if (type == "dielectric" || type == "thindielectric" || type == "roughdielectric") { // "Thin material are ignored / Impossible to do rough glass"
        auto kr = mts_color(cx, *b, "specularReflectance", Color{1, 1, 1}), kt = mts_color(cx, *b, "specularTransmittance", Color{1, 1, 1});
        Material m = Material::glass(kr.first, kt.first, xfloat(*b, "intIOR", 1.5046f), xfloat(*b, "extIOR", 1.000277f));
        m.m.ks_texture = kr.second, m.m.kt_texture = kt.second;
        return m;
    }
    if (type == "plastic" || type == "roughplastic") {
        auto ks = mts_color(cx, *b, "specularReflectance", Color{1, 1, 1}), kd = mts_color(cx, *b, "diffuseReflectance", Color{0.5f, 0.5f, 0.5f});
        float alpha;
        const uint32_t mf = mts_distribution(*b, type == "roughplastic", &alpha);
        return textured(Material::substrate(kd.first, ks.first, mf, alpha), kd.second, ks.second);
    }
    if (type == "conductor" || type == "roughconductor") {
        auto ks = mts_color(cx, *b, "specularReflectance", Color{1, 1, 1});
        const float ext = xfloat(*b, "extEta", 1.000277f);
        Color eta = xrgb(*b, "eta", Color{0.2004f, 0.9240f, 1.1022f}), k = xrgb(*b, "k", Color{3.9129f, 2.4528f, 2.1421f}); // copper, as for pbrt "metal"
        float alpha;
        const uint32_t mf = mts_distribution(*b, type == "roughconductor", &alpha);
        return textured(Material::metal(ks.first, Color{eta.r / ext, eta.g / ext, eta.b / ext}, Color{k.r / ext, k.g / ext, k.b / ext}, mf, alpha), 0u, ks.second);
    }
    return fallback; // `_ => None` -> BSDFDiffuse 0.8 (mod.rs:599-611)
}
// Mitsuba 0.5 / 0.6 ".serialized" mesh (mitsuba_rs::serialized::read_serialized, crate not vendored: the published format, UNPINNED):
//   [u16 0x041C][u16 version 3 | 4][zlib stream: u32 flags, (v4) UTF-8 name + NUL, u64 vertices, u64 triangles, positions 3 n,
//   normals 3 n if flags & 1, texcoords 2 n if flags & 2, colours 3 n if flags & 8 (skipped), indices 3 t (u32; u64 above 2^32 - 1 vertices)],
//   reals f32 (flag 0x1000) or f64 (0x2000, narrowed).  A file may hold several shapes: its last u32 is their count, preceded by the
//   table of their start offsets (u64 in version 4, u32 in version 3); shape 0 starts the file.
void read_serialized(const std::string &filename, uint32_t shape_index, RawShape &out, std::string *name_out) {
    const std::string file = read_file(filename);
    const unsigned char *fp = reinterpret_cast<const unsigned char *>(file.data());
    auto le = [&](size_t at, int bytes) {
        if (at + (size_t)bytes > file.size()) throw Error("serialized: truncated " + filename);
        uint64_t v = 0;
        for (int k = bytes - 1; k >= 0; k--) v = (v << 8) | fp[at + k];
        return v;
    };
    if (le(0, 2) != 0x041Cu) throw Error("serialized: bad magic in " + filename);
    size_t start = 0;
    if (shape_index > 0) {
        const uint64_t ver0 = le(2, 2), count = le(file.size() - 4, 4), osz = ver0 == 4 ? 8 : 4;
        if (shape_index >= count || file.size() < 4 + osz * count) throw Error("serialized: shape index out of range in " + filename);
        start = (size_t)le(file.size() - 4 - (size_t)(osz * (count - shape_index)), (int)osz);
        if (le(start, 2) != 0x041Cu) throw Error("serialized: bad shape offset in " + filename);
    }
    const uint64_t version = le(start + 2, 2);
    if (version != 3 && version != 4) throw Error("serialized: unknown version in " + filename);
    std::vector<unsigned char> data;
    {
        z_stream zs{};
        if (inflateInit(&zs) != Z_OK) throw Error("serialized: zlib");
        zs.next_in = const_cast<Bytef *>(fp + start + 4);
        zs.avail_in = (uInt)std::min<size_t>(file.size() - start - 4, 0xffffffffu);
        unsigned char buf[1 << 16];
        int rc = Z_OK;
        while (rc != Z_STREAM_END) {
            zs.next_out = buf, zs.avail_out = sizeof(buf);
            rc = inflate(&zs, Z_NO_FLUSH);
            if (rc != Z_OK && rc != Z_STREAM_END) {
                inflateEnd(&zs);
                throw Error("serialized: corrupt deflate stream in " + filename);
            }
            data.insert(data.end(), buf, buf + (sizeof(buf) - zs.avail_out));
            if (rc == Z_OK && zs.avail_in == 0 && zs.avail_out != 0) {
                inflateEnd(&zs);
                throw Error("serialized: truncated deflate stream in " + filename);
            }
        }
        inflateEnd(&zs);
    }
    size_t pos = 0;
    auto need = [&](size_t n) {
        if (n > data.size() - pos) throw Error("serialized: short mesh record in " + filename);
    };
    auto get = [&](int bytes) {
        need((size_t)bytes);
        uint64_t v = 0;
        for (int k = bytes - 1; k >= 0; k--) v = (v << 8) | data[pos + k];
        pos += (size_t)bytes;
        return v;
    };
    const uint32_t flags = (uint32_t)get(4);
    std::string name;
    if (version == 4) {
        while (true) {
            need(1);
            const char c = (char)data[pos++];
            if (!c) break;
            name.push_back(c);
        }
    }
    if (name_out) *name_out = name;
    const uint64_t nv = get(8), nt = get(8);
    const bool dbl = (flags & 0x2000u) != 0;
    if (!dbl && !(flags & 0x1000u)) throw Error("serialized: neither single nor double precision flagged in " + filename);
    if (nv == 0 || nt == 0 || nv > (1ull << 31) || nt > (1ull << 31)) throw Error("serialized: bad counts in " + filename);
    auto reals = [&](std::vector<float> *dst, size_t n) {
        const size_t sz = dbl ? 8 : 4;
        if (n > (data.size() - pos) / sz) throw Error("serialized: short mesh record in " + filename);
        if (dst) dst->resize(n);
        for (size_t i = 0; i < n; i++, pos += sz) {
            if (!dst) continue;
            if (dbl) {
                double d;
                std::memcpy(&d, &data[pos], 8);
                (*dst)[i] = (float)d;
            } else std::memcpy(&(*dst)[i], &data[pos], 4);
        }
    };
    reals(&out.P, (size_t)nv * 3);
    if (flags & 0x0001u) reals(&out.N, (size_t)nv * 3);
    if (flags & 0x0002u) reals(&out.uv, (size_t)nv * 2);
    if (flags & 0x0008u) reals(nullptr, (size_t)nv * 3);
    const int isz = nv > 0xffffffffull ? 8 : 4;
    out.idx.resize((size_t)nt * 3);
    for (size_t i = 0; i < out.idx.size(); i++) {
        const uint64_t v = get(isz);
        if (v >= nv) throw Error("serialized: vertex index out of range in " + filename);
        out.idx[i] = (uint32_t)v;
    }
}
// Wavefront OBJ, triangulated by fanning; one vertex per distinct (v, vt, vn) corner (tobj is not vendored: unpinned)
void read_obj(const std::string &filename, RawShape &out) {
    const std::string text = read_file(filename);
    std::vector<float> P, T, N;
    std::map<std::array<int, 3>, uint32_t> corner_ids;
    std::istringstream in(text);
    std::string line;
    bool any_n = false, any_t = false;
    std::vector<std::array<int, 3>> corners;
    while (std::getline(in, line)) {
        std::istringstream ls(line);
        std::string kw;
        if (!(ls >> kw)) continue;
        if (kw == "v" || kw == "vn") {
            float x, y, z;
            if (!(ls >> x >> y >> z)) throw Error("obj: bad " + kw + " line");
            auto &dst = kw == "v" ? P : N;
            dst.insert(dst.end(), {x, y, z});
        } else if (kw == "vt") {
            float u, v = 0.0f;
            if (!(ls >> u)) throw Error("obj: bad vt line");
            ls >> v;
            T.insert(T.end(), {u, v});
        } else if (kw == "f") {
            std::vector<uint32_t> face;
            std::string tok;
            while (ls >> tok) {
                std::array<int, 3> c{0, 0, 0};
                size_t a = tok.find('/');
                c[0] = std::atoi(tok.substr(0, a).c_str());
                if (a != std::string::npos) {
                    size_t b2 = tok.find('/', a + 1);
                    const std::string st = tok.substr(a + 1, b2 == std::string::npos ? std::string::npos : b2 - a - 1);
                    if (!st.empty()) c[1] = std::atoi(st.c_str());
                    if (b2 != std::string::npos) c[2] = std::atoi(tok.substr(b2 + 1).c_str());
                }
                if (c[0] < 0) c[0] = (int)(P.size() / 3) + 1 + c[0];
                if (c[1] < 0) c[1] = (int)(T.size() / 2) + 1 + c[1];
                if (c[2] < 0) c[2] = (int)(N.size() / 3) + 1 + c[2];
                if (c[0] <= 0 || (size_t)c[0] > P.size() / 3 || (size_t)c[1] > T.size() / 2 || (size_t)c[2] > N.size() / 3) throw Error("obj: index out of range");
                any_t = any_t || c[1] != 0, any_n = any_n || c[2] != 0;
                auto it = corner_ids.find(c);
                if (it == corner_ids.end()) {
                    it = corner_ids.emplace(c, (uint32_t)corners.size()).first;
                    corners.push_back(c);
                }
                face.push_back(it->second);
            }
            if (face.size() < 3) throw Error("obj: face with fewer than 3 vertices");
            for (size_t k = 1; k + 1 < face.size(); k++) out.idx.insert(out.idx.end(), {face[0], face[k], face[k + 1]});
        }
    }
    for (auto &c : corners) {
        out.P.insert(out.P.end(), {P[3 * (c[0] - 1)], P[3 * (c[0] - 1) + 1], P[3 * (c[0] - 1) + 2]});
        if (any_t) {
            if (c[1]) out.uv.insert(out.uv.end(), {T[2 * (c[1] - 1)], T[2 * (c[1] - 1) + 1]});
            else out.uv.insert(out.uv.end(), {0.0f, 0.0f});
        }
        if (any_n) {
            if (c[2]) out.N.insert(out.N.end(), {N[3 * (c[2] - 1)], N[3 * (c[2] - 1) + 1], N[3 * (c[2] - 1) + 2]});
            else out.N.insert(out.N.end(), {0.0f, 0.0f, 0.0f});
        }
    }
}
} // namespace

Scene MTSSceneLoader::load(const std::string &filename, bool use_shading_normal) const {
    size_t slash = filename.find_last_of('/');
    return load_string(read_file(filename), use_shading_normal, slash == std::string::npos ? "" : filename.substr(0, slash));
}
Scene MTSSceneLoader::load_string(const std::string &text, bool use_shading_normal, const std::string &base_dir) const {
    XParser xp(text);
    const XNode root = xp.element();
    if (root.tag != "scene") throw Error("xml: the root element must be <scene>");
    Scene scene;
    MtsCtx cx{&scene, base_dir, {}};
    std::vector<const XNode *> sensors, shapes_id, shapes_unnamed, emitters;
    for (auto &k : root.kids) {
        if ((k.tag == "bsdf" || k.tag == "texture") && k.attr.count("id")) cx.ids[k.get("id")] = &k;
        else if (k.tag == "sensor") sensors.push_back(&k);
        else if (k.tag == "shape") (k.attr.count("id") ? shapes_id : shapes_unnamed).push_back(&k);
        else if (k.tag == "emitter") emitters.push_back(&k);
        else if (k.tag == "medium") throw Error("xml: participating media are outside the hot-path scope");
    }
    if (sensors.size() != 1) throw Error("xml: exactly one sensor is expected (assert_eq!(mts.sensors.len(), 1), scene_loader.rs:328)");
    { // scene_loader.rs:327-338
        const XNode &se = *sensors[0];
        if (se.get("type") != "perspective") throw Error("xml: sensor type \"" + se.get("type") + "\" is not supported (perspective)");
        uint32_t w = 768, h = 576;
        for (auto &k : se.kids)
            if (k.tag == "film") w = (uint32_t)xfloat(k, "width", 768.0f), h = (uint32_t)xfloat(k, "height", 576.0f);
        const std::string axis = xstring(se, "fovAxis", "x");
        if (axis != "x" && axis != "y") throw Error("Unsupport Fov axis definition: " + axis); // :335
        const float fov = xfloat(se, "fov", 39.5978f); // Mitsuba's default focal length (50 mm on 36 mm film)
        scene.camera = Camera::create(w, h, axis == "x" ? Fov::X : Fov::Y, fov, xtransform(se.child("transform", "toWorld")), true);
    }
    auto emit = [&](const XNode &sh, RawShape &rs, bool keep_normals) { // option.{bsdf, emitter, to_world} + apply_transform (:341-376)
        const XNode *b = nullptr;
        for (auto &k : sh.kids) {
            if (k.tag == "ref" && !cx.ids.count(k.get("id"))) throw Error("xml: <ref id=\"" + k.get("id") + "\"> refers to nothing");
            if (k.tag == "bsdf" || (k.tag == "ref" && cx.ids[k.get("id")]->tag == "bsdf")) b = &k;
        }
        rs.bsdf = mts_bsdf(cx, b);
        for (auto &k : sh.kids)
            if (k.tag == "emitter") {
                if (k.get("type") != "area") throw Error("xml: a shape can only carry an area emitter");
                rs.is_light = true;
                rs.emission = xrgb(k, "radiance", Color{1, 1, 1});
            }
        if (!keep_normals) rs.N.clear();
        emit_mesh(scene, rs, xtransform(sh.child("transform", "toWorld")), true);
    };
    auto load_shape = [&](const XNode &sh) {
        const std::string type = sh.get("type");
        RawShape rs;
        const bool face_normal = xbool(sh, "faceNormals", false);
        if (type == "rectangle") { // :538-594
            rs.P = {-1, -1, 0, 1, -1, 0, 1, 1, 0, -1, 1, 0};
            rs.uv = {0, 0, 1, 0, 1, 1, 0, 1};
            rs.N = {0, 0, 1, 0, 0, 1, 0, 0, 1, 0, 0, 1};
            rs.idx = {0, 1, 2, 2, 3, 0};
            emit(sh, rs, true);
            scene.meshes.back()->name = "rectangle";
        } else if (type == "sphere") { // :596-665: 32 x 32 vertices, f32 sin / cos of the host (LIBM: the loader is host code in the reference too)
            Vec3 c{0, 0, 0};
            if (const XNode *p = sh.child("point", "center")) c = Vec3{std::strtof(p->get("x", "0").c_str(), nullptr), std::strtof(p->get("y", "0").c_str(), nullptr), std::strtof(p->get("z", "0").c_str(), nullptr)};
            const float radius = xfloat(sh, "radius", 1.0f);
            const int NB = 32;
            const float PI_F = 3.14159265358979323846f;
            for (int i = 0; i < NB; i++) {
                const float theta = (float)i / (float)(NB - 1) * PI_F;
                for (int j = 0; j < NB; j++) {
                    const float phi = (float)j / (float)(NB - 1) * 2.0f * PI_F;
                    const float x = radius * std::sin(theta) * std::cos(phi), y = radius * std::sin(theta) * std::sin(phi), z = radius * std::cos(theta);
                    rs.P.insert(rs.P.end(), {x + c.x, y + c.y, z + c.z});
                    const float il = 1.0f / std::sqrt(x * x + y * y + z * z); // normalize = v * (1 / |v|)
                    rs.N.insert(rs.N.end(), {x * il, y * il, z * il});
                    rs.uv.insert(rs.uv.end(), {theta / PI_F, phi / (2.0f * PI_F)});
                }
            }
            for (uint32_t i = 0; i + 1 < (uint32_t)NB; i++)
                for (uint32_t j = 0; j + 1 < (uint32_t)NB; j++) {
                    const uint32_t i0 = i * NB + j, i1 = i0 + 1, i2 = (i + 1) * NB + j + 1, i3 = (i + 1) * NB + j;
                    rs.idx.insert(rs.idx.end(), {i0, i1, i2, i2, i3, i0});
                }
            emit(sh, rs, true);
            scene.meshes.back()->name = "rectangle"; // (sic, :628)
        } else if (type == "ply" || type == "obj") {
            std::string fn = xstring(sh, "filename", "");
            if (fn.empty()) throw Error("xml: shape \"" + type + "\" needs a filename");
            if (fn[0] != '/' && !base_dir.empty()) fn = base_dir + "/" + fn;
            if (type == "ply") read_ply(fn, rs);
            else {
                read_obj(fn, rs);
                if (xbool(sh, "flipTexCoords", true)) // "Only v coordinate" (:482-491)
                    for (size_t k = 1; k < rs.uv.size(); k += 2) rs.uv[k] = 1.0f - rs.uv[k];
            }
            emit(sh, rs, !face_normal && (type == "obj" || use_shading_normal)); // ply: `face_normal || !use_shading_normal` -> None (:399-403)
        } else if (type == "serialized") { // :499-538: normals unless faceNormals (use_shading_normal is not consulted), texcoords as stored
            std::string fn = xstring(sh, "filename", "");
            if (fn.empty()) throw Error("xml: shape \"serialized\" needs a filename");
            if (fn[0] != '/' && !base_dir.empty()) fn = base_dir + "/" + fn;
            std::string mesh_name;
            read_serialized(fn, (uint32_t)xfloat(sh, "shapeIndex", 0.0f), rs, &mesh_name);
            emit(sh, rs, !face_normal);
            scene.meshes.back()->name = mesh_name; // (:502)
        } // anything else: "Ignoring shape" (:666-669)
    };
    for (const XNode *sh : shapes_id) load_shape(*sh);
    for (const XNode *sh : shapes_unnamed) load_shape(*sh);
    for (const XNode *e : emitters) { // :676-724
        const std::string type = e->get("type");
        if (type == "point") {
            Vec3 p{0, 0, 0};
            if (const XNode *q = e->child("point", "position")) p = Vec3{std::strtof(q->get("x", "0").c_str(), nullptr), std::strtof(q->get("y", "0").c_str(), nullptr), std::strtof(q->get("z", "0").c_str(), nullptr)};
            p = xtransform(e->child("transform", "toWorld")).transform_point(p);
            Color I = xrgb(*e, "intensity", Color{1, 1, 1});
            scene.add_point_light(I, p.x, p.y, p.z);
        } else if (type == "pointnormal" || type == "PointNormal")
            throw Error("xml: the PointNormal emitter cannot be light-sampled (PointNormalEmitter::direct_sample is todo!() in the reference, emitter.rs:262-264)");
        // anything else: "Ignoring emitter" (:720-722)
    }
    if (scene.meshes.empty()) throw Error("xml: the scene has no shape this loader reads");
    return scene;
}

// ------------------------------------------------------------------------------------------
// SceneLoaderManager, src/scene_loader.rs:21-58
// ------------------------------------------------------------------------------------------
SceneLoaderManager::SceneLoaderManager() {
    loader["pbrt"] = std::make_shared<PBRTSceneLoader>();
    loader["json"] = std::make_shared<JSONSceneLoader>();
    loader["xml"] = std::make_shared<MTSSceneLoader>(); // scene_loader.rs:52-56 (feature "mitsuba")
}
Scene SceneLoaderManager::load(const std::string &filename, bool use_shading_normal) const {
    size_t dot = filename.find_last_of('.');
    size_t slash = filename.find_last_of('/');
    if (dot == std::string::npos || (slash != std::string::npos && dot < slash))
        throw Error("No file extension provided"); // scene_loader.rs:34
    std::string ext = filename.substr(dot + 1);
    auto it = loader.find(ext);
    if (it == loader.end())
        throw Error("Impossible to found scene loader for " + ext + " extension"); // scene_loader.rs:40-43
    return it->second->load(filename, use_shading_normal);
}

} // namespace rlh
