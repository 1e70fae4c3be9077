/* c_abi_smoke.c -- a plain C program against include/rl_b200.h and librl_b200.so (no Python, no C++): the smallest user of the
 * drop-in boundary.  It describes a scene by hand (a floor quad under a square lamp; the camera matrices come from the host
 * library's Camera::new, rlh_camera_create), renders 16x16 pixels with `path` twice (two passes of an averaging wrapper: the
 * second pass continues the sample sequence) and once with `direct`, and checks what does not need an oracle: status codes,
 * finite non-negative radiance, light reaches the floor, the two passes differ, trace/visible agree with the geometry.
 *   gcc -Iinclude tests/c_abi_smoke.c -Lrustlight_b200 -lrl_b200 -lrl_host -Wl,-rpath,rustlight_b200 -lm
 * Exit code 0 = ok; 3 = no CUDA device (rl_create said so loudly: this library has no CPU path). */
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "rl_b200.h"

int rlh_camera_create(uint32_t w, uint32_t h, int fov_axis, float fov_deg, const float to_world[16], int flip, float out_sample_to_camera[16],
                      float out_camera_to_sample[16]);

#define FAIL() do { fprintf(stderr, "c_abi_smoke: check failed at line %d\n", __LINE__); return 1; } while (0)
#define CHECK(call)                                                                    \
    do {                                                                               \
        int rc_ = (call);                                                              \
        if (rc_ != RL_OK) {                                                            \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc_, rl_last_error(ctx));         \
            return 1;                                                                  \
        }                                                                              \
    } while (0)

int main(void) {
    if (rl_abi_version() != RL_B200_ABI_VERSION) {
        fprintf(stderr, "header / library ABI mismatch: %d vs %d\n", RL_B200_ABI_VERSION, rl_abi_version());
        return 1;
    }
    rl_ctx *ctx = NULL;
    int rc = rl_create(&ctx, 0, 1, 0, NULL);
    if (rc == RL_ERR_CUDA) {
        fprintf(stderr, "rl_create: %s\n", rl_last_error(NULL));
        return 3;
    }
    if (rc != RL_OK) FAIL();

    /* floor y = 0 (4 x 4), lamp y = 2 (1 x 1, facing down) */
    static const float floor_p[] = {-2, 0, -2, 2, 0, -2, 2, 0, 2, -2, 0, 2};
    static const float lamp_p[] = {-0.5f, 2, -0.5f, 0.5f, 2, -0.5f, 0.5f, 2, 0.5f, -0.5f, 2, 0.5f};
    static const uint32_t floor_i[] = {0, 2, 1, 0, 3, 2}; /* normal +y */
    static const uint32_t lamp_i[] = {0, 1, 2, 0, 2, 3};  /* normal -y */
    rl_mesh_desc meshes[2];
    memset(meshes, 0, sizeof(meshes));
    meshes[0].P = floor_p, meshes[0].nverts = 4, meshes[0].idx = floor_i, meshes[0].ntris = 2;
    meshes[0].mat.kind = RL_BSDF_DIFFUSE, meshes[0].mat.kd[0] = 0.5f, meshes[0].mat.kd[1] = 0.6f, meshes[0].mat.kd[2] = 0.7f;
    meshes[1].P = lamp_p, meshes[1].nverts = 4, meshes[1].idx = lamp_i, meshes[1].ntris = 2;
    meshes[1].mat.kind = RL_BSDF_DIFFUSE, meshes[1].emission_kind = 1;
    meshes[1].emission[0] = meshes[1].emission[1] = meshes[1].emission[2] = 10.0f;
    rl_scene_desc desc;
    memset(&desc, 0, sizeof(desc));
    desc.nmeshes = 2, desc.meshes = meshes;
    desc.camera.width = 16, desc.camera.height = 16;
    /* camera at (0, 1, 5); Camera::new looks along +z of camera space (camera.rs:50: "undo gluPerspective"), so camera z maps to world -z.
     * Column-major camera-to-world. */
    const float to_world[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, -1, 0, 0, 1, 5, 1};
    memcpy(desc.camera.to_world, to_world, sizeof(to_world));
    if (rlh_camera_create(16, 16, 0, 40.0f, to_world, 0, desc.camera.sample_to_camera, NULL) != 0) FAIL();

    rl_scene *scene = NULL;
    CHECK(rl_scene_create(ctx, &desc, &scene));
    rl_bvh_info info;
    CHECK(rl_scene_bvh_info(ctx, scene, &info));
    if (info.ntris != 4) FAIL();

    rl_integrator_desc path;
    memset(&path, 0, sizeof(path));
    path.kind = RL_INTEGRATOR_PATH, path.min_depth = 0, path.max_depth = -1, path.rr_depth = 0, path.strategy = RL_STRATEGY_ALL;
    path.nb_bsdf_samples = path.nb_light_samples = 1, path.ao_max_distance = -1.0f;
    rl_render_opts opts;
    memset(&opts, 0, sizeof(opts));
    opts.struct_size = sizeof(opts), opts.spp = 32, opts.seed = 7, opts.sampler_mode = RL_SAMPLER_COUNTER, opts.material_sort = 2;
    static float img[2][16 * 16 * 3], dimg[16 * 16 * 3];
    rl_stats st;
    CHECK(rl_render(ctx, scene, &path, &opts, img[0], &st));
    if (st.samples != 16u * 16u * 32u || st.segments < st.samples) FAIL();
    opts.sample_offset = 32; /* pass 2 of an averaging wrapper */
    CHECK(rl_render(ctx, scene, &path, &opts, img[1], &st));
    double sum = 0.0, diff = 0.0;
    for (int i = 0; i < 16 * 16 * 3; i++) {
        if (!(img[0][i] >= 0.0f) || !isfinite(img[0][i])) FAIL();
        sum += img[0][i], diff += fabs((double)img[0][i] - (double)img[1][i]);
    }
    if (!(sum > 1.0) || !(diff > 0.0)) {
        fprintf(stderr, "path: sum %g, pass difference %g\n", sum, diff);
        return 1;
    }
    rl_integrator_desc direct = path;
    direct.kind = RL_INTEGRATOR_DIRECT;
    opts.sample_offset = 0;
    CHECK(rl_render(ctx, scene, &direct, &opts, dimg, &st));
    if (st.segments > 2 * st.samples) FAIL();

    /* Acceleration::trace / visible: straight down onto the floor from y = 1; floor point <-> lamp centre */
    const float o[3] = {0.25f, 1.0f, 0.25f}, d[3] = {0, -1, 0};
    uint32_t prim = 0;
    float tuv[3];
    CHECK(rl_trace(ctx, scene, 1, o, d, &prim, tuv));
    if (prim > 1 || fabsf(tuv[0] - 1.0f) > 1e-6f) FAIL();
    const float p0[6] = {0.25f, 0.0f, 0.25f, 0.25f, 0.0f, 0.25f}, p1[6] = {0.0f, 2.0f, 0.0f, 0.0f, -3.0f, 0.0f};
    uint8_t vis[2];
    CHECK(rl_visible(ctx, scene, 2, p0, p1, vis));
    if (vis[0] != 1) FAIL();

    /* errors are status codes, never aborts */
    opts.spp = 0;
    if (rl_render(ctx, scene, &path, &opts, img[0], &st) != RL_ERR_INVALID) FAIL();
    rl_scene_destroy(ctx, scene);
    rl_destroy(ctx);
    printf("c_abi_smoke ok: path mean %.5f\n", sum / (16 * 16 * 3));
    return 0;
}
