// rl_integrators.hpp -- C++ mirror of the reference's integrator objects on top of the C ABI.
//   Integrator / IntegratorType::compute      src/integrators/mod.rs:219-233, 274-338
//   IntegratorPathTracing                      src/integrators/explicit/path.rs:14-20, 187-196
//   IntegratorDirect                           src/integrators/direct.rs:5-19
//   IntegratorAverage (`-a`)                   src/integrators/avg.rs:5-131
//   IntegratorEqualTime (`-e`)                 src/integrators/equal_time.rs:4-66
// `compute` keeps the reference's meaning (one frame of scene.nb_samples samples per pixel);
// the sampler object carries the seed and the number of passes already drawn from it, which
// plays the role of the reference's stateful `&mut dyn Sampler`.
#pragma once
#include <chrono>
#include <cstdio>
#include <memory>

#include "rl_b200.h"
#include "rl_host.hpp"

namespace rlh {

struct IndependentSampler { // samplers/independent.rs (`-r independent:<seed>`, cli.rs:886-890)
    uint64_t seed = 0;
    uint32_t passes = 0; // frames rendered so far with this sampler
};

// Owns the rl_ctx and the device-resident scene (what BVHAccel::new + build_emitters prepare).
class Device {
  public:
    explicit Device(int device = 0) {
        if (rl_create(&ctx_, device, 1, 0, nullptr) != RL_OK) throw Error(std::string("rl_create: ") + rl_last_error(nullptr));
    }
    ~Device() {
        if (scene_) rl_scene_destroy(ctx_, scene_);
        if (ctx_) rl_destroy(ctx_);
    }
    void upload(Scene &scene) {
        if (scene_) rl_scene_destroy(ctx_, scene_);
        scene_ = nullptr;
        if (rl_scene_create(ctx_, scene.desc(), &scene_) != RL_OK) throw Error(std::string("rl_scene_create: ") + rl_last_error(ctx_));
    }
    rl_ctx *ctx() const { return ctx_; }
    rl_scene *scene() const { return scene_; }

  private:
    rl_ctx *ctx_ = nullptr;
    rl_scene *scene_ = nullptr;
};

struct Integrator {
    virtual ~Integrator() = default;
    virtual BufferCollection compute(IndependentSampler &sampler, Device &dev, Scene &scene) = 0;
    virtual bool averaging() const { return true; }
};

inline BufferCollection render_primal(Device &dev, Scene &scene, const rl_integrator_desc &integ, IndependentSampler &sampler, rl_stats *stats_out = nullptr) {
    rl_render_opts opts{};
    opts.struct_size = sizeof(opts);
    opts.spp = (uint32_t)scene.nb_samples;
    opts.seed = sampler.seed;
    opts.sampler_mode = RL_SAMPLER_COUNTER;
    opts.material_sort = 2; // auto: on when the scene mixes BSDF kinds
    opts.sample_offset = sampler.passes * (uint32_t)scene.nb_samples;
    Bitmap bmp;
    bmp.size_x = scene.camera.img_x, bmp.size_y = scene.camera.img_y;
    bmp.colors.assign((size_t)3 * bmp.size_x * bmp.size_y, 0.0f);
    rl_stats st{};
    if (rl_render(dev.ctx(), dev.scene(), &integ, &opts, bmp.colors.data(), &st) != RL_OK) throw Error(std::string("rl_render: ") + rl_last_error(dev.ctx()));
    sampler.passes++;
    if (stats_out) *stats_out = st;
    BufferCollection bc;
    bc.values["primal"] = std::move(bmp);
    return bc;
}

enum class IntegratorPathTracingStrategies { All, BSDF, Emitter };
struct IntegratorPathTracing : Integrator {
    std::optional<uint32_t> min_depth = 0, max_depth, rr_depth = 0; // CLI defaults, cli.rs:54-61
    IntegratorPathTracingStrategies strategy = IntegratorPathTracingStrategies::All;
    bool single_scattering = false;
    rl_stats last_stats{};
    BufferCollection compute(IndependentSampler &sampler, Device &dev, Scene &scene) override {
        rl_integrator_desc d{};
        d.kind = RL_INTEGRATOR_PATH;
        d.min_depth = min_depth ? (int32_t)*min_depth : -1;
        d.max_depth = max_depth ? (int32_t)*max_depth : -1;
        d.rr_depth = rr_depth ? (int32_t)*rr_depth : -1;
        d.strategy = (uint32_t)strategy;
        d.single_scattering = single_scattering ? 1u : 0u;
        d.nb_bsdf_samples = d.nb_light_samples = 1;
        return render_primal(dev, scene, d, sampler, &last_stats);
    }
};
struct IntegratorDirect : Integrator {
    uint32_t nb_bsdf_samples = 1, nb_light_samples = 1; // cli.rs:157-160
    rl_stats last_stats{};
    BufferCollection compute(IndependentSampler &sampler, Device &dev, Scene &scene) override {
        rl_integrator_desc d{};
        d.kind = RL_INTEGRATOR_DIRECT;
        d.min_depth = 0, d.max_depth = -1, d.rr_depth = 0;
        d.nb_bsdf_samples = nb_bsdf_samples, d.nb_light_samples = nb_light_samples;
        return render_primal(dev, scene, d, sampler, &last_stats);
    }
};

// ao.rs:4-72 (`ao -d <distance|inf> -n`, cli.rs:150-155)
struct IntegratorAO : Integrator {
    std::optional<float> max_distance = 1.0f;
    bool normal_correction = false;
    rl_stats last_stats{};
    BufferCollection compute(IndependentSampler &sampler, Device &dev, Scene &scene) override {
        rl_integrator_desc d{};
        d.kind = RL_INTEGRATOR_AO;
        d.min_depth = 0, d.max_depth = -1, d.rr_depth = 0;
        d.nb_bsdf_samples = 1, d.nb_light_samples = 0;
        d.ao_max_distance = max_distance ? *max_distance : -1.0f;
        d.ao_normal_correction = normal_correction ? 1u : 0u;
        return render_primal(dev, scene, d, sampler, &last_stats);
    }
};

inline void bitmap_scale(Bitmap &b, float f) { // Bitmap::scale, structure.rs:423-425
    for (float &c : b.colors) c *= f;
}
inline void bitmap_accumulate(Bitmap &b, const Bitmap &o) { // accumulate_bitmap, structure.rs:406-415
    for (size_t i = 0; i < b.colors.size(); i++) b.colors[i] += o.colors[i];
}

// avg.rs:11-131
struct IntegratorAverage : Integrator {
    std::optional<size_t> time_out; // seconds
    std::unique_ptr<Integrator> integrator;
    bool dump_all = true;
    std::optional<size_t> max_iterations; // extension for tests: stop after this many passes
    BufferCollection compute(IndependentSampler &sampler, Device &dev, Scene &scene) override {
        if (!dump_all && !time_out && !max_iterations) throw Error("Impossible to have infinite approach and not dumping all images");
        const std::string &out = scene.output_img_path;
        size_t dot = out.find_last_of('.');
        if (dot == std::string::npos) throw Error("No file extension provided");
        std::string base = out.substr(0, dot), ext = out.substr(dot + 1);
        FILE *csv = dump_all ? std::fopen((base + "_time.csv").c_str(), "w") : nullptr;
        BufferCollection bitmap;
        size_t iteration = 1;
        double time_rendering = 0.0;
        for (;;) {
            auto start = std::chrono::steady_clock::now();
            BufferCollection nb = integrator->compute(sampler, dev, scene);
            if (iteration == 1) bitmap = std::move(nb);
            else if (integrator->averaging()) {
                Bitmap &b = bitmap.values["primal"];
                bitmap_scale(b, (float)iteration);
                bitmap_accumulate(b, nb.values["primal"]);
                bitmap_scale(b, 1.0f / (float)(iteration + 1));
            } else bitmap = std::move(nb);
            time_rendering += std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
            if (dump_all) bitmap.save("primal", base + "_" + std::to_string(iteration) + "." + ext);
            if (csv) {
                std::fprintf(csv, "%llu.%u,\n", (unsigned long long)time_rendering, (unsigned)((time_rendering - (double)(unsigned long long)time_rendering) * 1000.0));
                std::fflush(csv);
            }
            if (time_out && (size_t)time_rendering >= *time_out) break;
            if (max_iterations && iteration >= *max_iterations) break;
            iteration++;
        }
        if (csv) std::fclose(csv);
        return bitmap;
    }
};

// equal_time.rs:9-66
struct IntegratorEqualTime : Integrator {
    uint64_t target_time_ms = 0;
    std::unique_ptr<Integrator> integrator;
    size_t iterations_done = 0;
    BufferCollection compute(IndependentSampler &sampler, Device &dev, Scene &scene) override {
        BufferCollection bitmap;
        size_t iteration = 1;
        double ms = 0.0;
        for (;;) {
            auto start = std::chrono::steady_clock::now();
            BufferCollection nb = integrator->compute(sampler, dev, scene);
            if (iteration == 1) bitmap = std::move(nb);
            else if (integrator->averaging()) bitmap_accumulate(bitmap.values["primal"], nb.values["primal"]);
            else bitmap = std::move(nb);
            ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - start).count();
            if ((uint64_t)ms >= target_time_ms) break;
            iteration++;
        }
        bitmap_scale(bitmap.values["primal"], 1.0f / (float)iteration);
        iterations_done = iteration;
        return bitmap;
    }
};

} // namespace rlh
