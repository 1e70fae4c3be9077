// cli.cpp -- `rustlight-b200`: the reference's command line (examples/cli.rs:106-275, wiring 277-924) for the
// integrators that run on the GPU.
//
//   rustlight-b200 [-n spp] [-a passes|inf|<secs>s] [-e secs] [-r independent[:seed]] [-s scale] -o out.pfm
//                  [-t threads] [-m density] [-l logfile] [-x no-shading]... <scene.{pbrt,json}>
//                  path   [-m max_depth|inf] [-n min_depth] [-r rr_depth|inf] [-x] [-s all|bsdf|emitter]
//                | direct [-b nb_bsdf_samples] [-l nb_light_samples]
//                | ao     [-d distance|inf] [-n]
//
// Differences from the reference, all forced by scope: only `path`, `direct` and `ao`; `-m` must be 0
// (no medium); `-x ats|hvs-light|texture-light|no-shading` as in the reference (texture-light reads butterfly.jpg like the reference, or butterfly.png / .pfm); `-t` is accepted and ignored (the GPU replaces
// the Rayon pool); output is .pfm or .png (gamma 2.2, 8 bit, as Bitmap::save_ldr_image).  `-a N` averages N passes (the reference's argument is a time-out in
// seconds or `inf`; both spellings are accepted: `-a 30s` / `-a inf` / `-a 8`).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "rl_integrators.hpp"

using namespace rlh;

static std::optional<uint32_t> match_infinity(const std::string &s) { // cli.rs:28-39
    if (s == "inf") return std::nullopt;
    return (uint32_t)std::stoul(s);
}
[[noreturn]] static void usage(const char *msg) {
    std::fprintf(stderr, "error: %s\nusage: rustlight-b200 [-n N] [-a A] [-e SECS] [-r independent[:SEED]] [-s SCALE] -o OUT.pfm|OUT.png [-x no-shading|ats] SCENE "
                         "(path [-m MAX] [-n MIN] [-r RR] [-x] [-s all|bsdf|emitter] | direct [-b NB] [-l NL] | ao [-d DIST|inf] [-n])\n", msg);
    std::exit(2);
}

int main(int argc, char **argv) {
    std::vector<std::string> a(argv + 1, argv + argc);
    size_t nbsamples = 1;
    std::optional<std::string> average, equal_time;
    std::string rng = "independent", output, scene_path, medium = "0.0";
    float scale_image = 1.0f;
    bool shading_normals = true, use_ats = false, hsv_lights = false, texture_lights = false;
    size_t i = 0;
    auto need = [&](const char *flag) -> std::string {
        if (i + 1 >= a.size()) usage((std::string("missing value for ") + flag).c_str());
        return a[++i];
    };
    std::string command;
    for (; i < a.size(); i++) {
        const std::string &t = a[i];
        if (t == "path" || t == "direct" || t == "ao") {
            command = t;
            i++;
            break;
        } else if (t == "-n" || t == "--nbsamples") nbsamples = std::stoul(need("-n"));
        else if (t == "-a" || t == "--average") average = need("-a");
        else if (t == "-t" || t == "--threads") need("-t");
        else if (t == "-r" || t == "--random-number-generator") rng = need("-r");
        else if (t == "-s" || t == "--scale-image") scale_image = std::stof(need("-s"));
        else if (t == "-e" || t == "--equal-time") equal_time = need("-e");
        else if (t == "-o" || t == "--output") output = need("-o");
        else if (t == "-m" || t == "--medium") medium = need("-m");
        else if (t == "-l" || t == "--log") need("-l");
        else if (t == "-x" || t == "--xtra-options") {
            std::string x = need("-x");
            if (x == "no-shading") shading_normals = false;
            else if (x == "ats") use_ats = true; // Scene::build_emitters(true), cli.rs:325, 432
            else if (x == "hvs-light") hsv_lights = true;         // EmissionType::HSV on every mesh light, cli.rs:410-421
            else if (x == "texture-light") texture_lights = true; // EmissionType::Texture { img: Bitmap::read("butterfly.jpg") }, cli.rs:421-427
            else usage(("-x " + x + " is outside the GPU path").c_str());
        } else if (t == "-h" || t == "--help") usage("help");
        else if (!t.empty() && t[0] == '-') usage(("unknown option " + t).c_str());
        else scene_path = t;
    }
    if (command.empty()) usage("missing subcommand (path | direct | ao)");
    if (scene_path.empty()) usage("missing scene file");
    if (output.empty()) usage("missing -o output");
    if (std::stof(medium) != 0.0f) usage("participating media are outside the GPU path (-m must be 0)");
    {
        const std::string ext = output.size() >= 4 ? output.substr(output.size() - 4) : std::string();
        if (ext != ".pfm" && ext != ".png") usage("output must be a .pfm or .png file (Bitmap::save, structure.rs:528-545; .exr needs the reference's optional openexr feature)");
    }

    std::unique_ptr<Integrator> integ;
    if (command == "path") {
        auto p = std::make_unique<IntegratorPathTracing>();
        for (; i < a.size(); i++) {
            const std::string &t = a[i];
            if (t == "-m" || t == "--max-depth") p->max_depth = match_infinity(need("-m"));
            else if (t == "-n" || t == "--min-depth") p->min_depth = match_infinity(need("-n"));
            else if (t == "-r" || t == "--rr-depth") p->rr_depth = match_infinity(need("-r"));
            else if (t == "-x" || t == "--single-scattering") p->single_scattering = true;
            else if (t == "-s" || t == "--strategy") {
                std::string s = need("-s");
                if (s == "all") p->strategy = IntegratorPathTracingStrategies::All;
                else if (s == "bsdf") p->strategy = IntegratorPathTracingStrategies::BSDF;
                else if (s == "emitter") p->strategy = IntegratorPathTracingStrategies::Emitter;
                else usage("invalid strategy: [all, bsdf, emitter]"); // cli.rs:536-541
            } else usage(("unknown path option " + t).c_str());
        }
        integ = std::move(p);
    } else if (command == "ao") { // cli.rs:150-155, 856-865
        auto p = std::make_unique<IntegratorAO>();
        for (; i < a.size(); i++) {
            const std::string &t = a[i];
            if (t == "-d" || t == "--distance") {
                std::string v = need("-d");
                if (v == "inf") p->max_distance = std::nullopt;
                else p->max_distance = std::stof(v);
            } else if (t == "-n" || t == "--normal-correction") p->normal_correction = true;
            else usage(("unknown ao option " + t).c_str());
        }
        integ = std::move(p);
    } else {
        auto d = std::make_unique<IntegratorDirect>();
        for (; i < a.size(); i++) {
            const std::string &t = a[i];
            if (t == "-b" || t == "--nb-bsdf-samples") d->nb_bsdf_samples = (uint32_t)std::stoul(need("-b"));
            else if (t == "-l" || t == "--nb-light-samples") d->nb_light_samples = (uint32_t)std::stoul(need("-l"));
            else usage(("unknown direct option " + t).c_str());
        }
        integ = std::move(d);
    }
    try {
        Scene scene = SceneLoaderManager().load(scene_path, shading_normals);
        scene.use_ats = use_ats;
        if (hsv_lights) scene.override_lights_hsv(); // (hvs wins when both are given, cli.rs:419)
        else if (texture_lights) {
            // the reference reads "butterfly.jpg" from the working directory (Bitmap::read, cli.rs:424); the same picture as .png / .pfm is
            // accepted too, and its absence is an error like the reference's panic
            uint32_t id = 0;
            for (const char *fn : {"butterfly.jpg", "butterfly.png", "butterfly.pfm"}) {
                if (std::ifstream(fn).good()) {
                    id = scene.add_texture(Texture::bitmap_file(fn));
                    break;
                }
            }
            if (!id) throw Error("-x texture-light: butterfly.jpg (or .png / .pfm) not found in the working directory");
            scene.override_lights_texture(id);
        }
        scene.nb_samples = nbsamples;
        scene.output_img_path = output;
        if (scale_image != 1.0f) {
            if (scale_image == 0.0f) usage("image scale must not be 0");
            scene.camera.scale_image(scale_image);
        }
        IndependentSampler sampler;
        if (rng.rfind("independent", 0) != 0) usage("only `-r independent[:seed]` runs on the GPU");
        if (rng.size() > 12 && rng[11] == ':') sampler.seed = std::stoull(rng.substr(12));
        // wrappers, cli.rs:892-916 (equal time first, then averaging)
        if (equal_time) {
            auto w = std::make_unique<IntegratorEqualTime>();
            w->target_time_ms = (uint64_t)(std::stod(*equal_time) * 1000.0);
            w->integrator = std::move(integ);
            integ = std::move(w);
        }
        if (average) {
            auto w = std::make_unique<IntegratorAverage>();
            const std::string &v = *average;
            if (v == "inf") w->time_out = std::nullopt;
            else if (v.back() == 's') w->time_out = (size_t)std::stoul(v.substr(0, v.size() - 1));
            else w->max_iterations = (size_t)std::stoul(v);
            w->integrator = std::move(integ);
            integ = std::move(w);
        }
        std::fprintf(stderr, "INFO rustlight_b200 - Build acceleration data structure...\n");
        Device dev(0);
        dev.upload(scene);
        std::fprintf(stderr, "INFO rustlight_b200 - Run Integrator...\n");
        auto start = std::chrono::steady_clock::now();
        BufferCollection img = integ->compute(sampler, dev, scene);
        auto ms = std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - start).count();
        std::fprintf(stderr, "INFO rustlight_b200 - Elapsed Integrator: %lld ms\n", (long long)ms); // integrators/mod.rs:334
        img.save("primal", output);
    } catch (const std::exception &e) {
        std::fprintf(stderr, "ERROR rustlight_b200 - %s\n", e.what());
        return 1;
    }
    return 0;
}
