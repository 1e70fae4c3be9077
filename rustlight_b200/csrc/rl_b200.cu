// rl_b200.cu -- implementation of include/rl_b200.h: context, scene upload + device LBVH build,
// and the wavefront render loop.  Everything that touches a pixel sample runs in the kernels
// of rl_kernels.cuh; there is no CPU path (calls fail with RL_ERR_CUDA when no device exists).
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <cmath>
#include <mutex>
#include <vector>

#include <cub/device/device_radix_sort.cuh>

#include "rl_b200.h"
#include "rl_flat_host.hpp"
#include "rl_kernels.cuh"
#include "rl_refbvh_host.hpp"
#include "rl_scene_host.hpp"
#include "rl_wide_host.hpp"

using namespace rl;

// ---- minimal NCCL surface, resolved at run time so the library has no link-time dependency -----
namespace {
typedef struct ncclComm *ncclComm_t;
struct NcclApi {
    void *handle = nullptr;
    int (*GetUniqueId)(void *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, char[128], int) = nullptr; // ncclUniqueId is passed by value (128-byte struct)
    int (*Reduce)(const void *, void *, size_t, int, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
struct NcclId {
    char internal[128];
};
NcclApi g_nccl;
bool load_nccl(std::string &err) {
    if (g_nccl.handle) return true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        g_nccl.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.handle) break;
    }
    if (!g_nccl.handle) {
        err = "libnccl.so.2 not found (dlopen)";
        return false;
    }
    g_nccl.GetUniqueId = (int (*)(void *))dlsym(g_nccl.handle, "ncclGetUniqueId");
    *(void **)&g_nccl.CommInitRank = dlsym(g_nccl.handle, "ncclCommInitRank");
    *(void **)&g_nccl.Reduce = dlsym(g_nccl.handle, "ncclReduce");
    *(void **)&g_nccl.CommDestroy = dlsym(g_nccl.handle, "ncclCommDestroy");
    *(void **)&g_nccl.GetErrorString = dlsym(g_nccl.handle, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.Reduce || !g_nccl.CommDestroy) {
        err = "libnccl.so.2 lacks the expected symbols";
        return false;
    }
    return true;
}
std::string g_create_error;
} // namespace

struct rl_ctx {
    int device = 0, nranks = 1, rank = 0, sm_count = 148;
    size_t mem_total = 0; // device memory (sizes the wavefront batches)
    cudaStream_t stream = nullptr;
    std::string err;
    int profiling = 0; // 0 off, 1 per-stage kernels, 2 events around the kernels of an untimed frame
    ncclComm_t comm = nullptr;
    // wavefront buffers (grown on demand)
    size_t cap_paths = 0, cap_shadow = 0, cap_lacc = 0; // ray/hit queues, shadow queues, radiance slots (records)
    float4 *ray_o[2] = {nullptr, nullptr}, *ray_d[2] = {nullptr, nullptr}, *state[2] = {nullptr, nullptr};
    float4 *hit = nullptr, *sh_a = nullptr, *sh_b = nullptr, *sh_c = nullptr, *lacc = nullptr;
    uint32_t *fix_trace = nullptr, *fix_shadow = nullptr; // queue indices of the rays / shadow segments k_fix_flat re-traces (sized like the queues)
    // per-image buffers
    size_t cap_pix = 0, cap_frame = 0;
    float4 *img_sum = nullptr;
    uint32_t *pixel_list = nullptr;
    float *frame = nullptr; // W*H*3 mean image of this rank (zeros outside its tiles)
    uint32_t pl_w = 0, pl_h = 0, pl_npix = 0;
    uint64_t pl_gen = 0; // bumped whenever the pixel list changes (keys the per-scene camera masks)
    uint32_t *d_counts = nullptr; // [cur/next ping-pong x2, shadow, pad] (direct integrator, rl_trace)
    uint32_t *d_hist = nullptr, *h_hist = nullptr; // queue length per wavefront iteration [0,kMaxIters) and shadow-queue length [kMaxIters, 2*kMaxIters)
    Counters *d_counters = nullptr;
    uint32_t *h_counts = nullptr; // pinned
    Counters *h_counters = nullptr;
    cudaEvent_t ev[8] = {};
    uint64_t launches = 0;
    // per-launch events of a profiled frame (rl_set_profiling) and the queue-length prediction of the device-decided schedule
    std::vector<cudaEvent_t> ev_pool;
    std::vector<uint8_t> ev_kind;
    size_t ev_used = 0;
    std::vector<double> pred_ratio; // queue length of iteration k / paths of the batch, last frame
    uint64_t pred_key = 0;
};

struct rl_scene {
    HostScene hs;
    SceneView sv{};
    float4 *d_flat = nullptr; // group table of small scenes (rl_flat_host.hpp), nullptr when absent
    float4 *d_quad_verts = nullptr; // vertices of the table's quads (k_camera_cull)
    float4 *d_ref_nodes = nullptr;  // the reference's own BVH (rl_refbvh_host.hpp): visit order of tied hits
    uint32_t *d_ref_prims = nullptr, *d_ref_up = nullptr;
    uint32_t *d_cam_masks = nullptr; // per block of 32 local pixels: quads its camera rays can see
    size_t cam_cap = 0;
    uint64_t cam_gen = 0; // ctx->pl_gen the masks were computed for (0 = never)
    FlatTable flat;
    float4 *d_trav = nullptr, *d_nodes = nullptr, *d_shade = nullptr, *d_verts = nullptr, *d_mats = nullptr, *d_emit_info = nullptr;
    float *d_emit_cdf = nullptr, *d_area_cdf = nullptr;
    float2 *d_uvs = nullptr;
    float4 *d_tex = nullptr, *d_texels = nullptr;
    bool emit_var = false; // a mesh light with EmissionType::HSV / Texture
    float *d_env_dist = nullptr; // Distribution2D of an environment texture
    float4 *d_ats_nodes = nullptr; // LightSamplerATS (rl_ats_host.hpp)
    uint32_t *d_ats_leaf = nullptr;
    uint32_t n_node_f4 = 0, n_trav_f4 = 0, n_ref_f4 = 0;
    size_t smem_bytes = 0;
    bool smem_ok = false;
    rl_bvh_info info{};
    // roots: the tree (coherent rays) and, for scenes of <= 64 triangles, the whole scene as one leaf
    int root_tree = 0, root_flat = 0, coherent_tree = 2;
    uint32_t n_bsdf_kinds = 1; // distinct rl_bsdf_kind values over the meshes (material sort: auto)
    uint32_t kind_mask = 1;    // bit k: some mesh has rl_bsdf_kind k (selects the k_shade specialisation)
    bool flat_ok = false; // group table present: incoherent rays use k_trace_flat / k_shadow_flat
    size_t smem_flat_bytes = 0;
    uint32_t scene_gen = 0; // distinguishes scenes that reuse an address (queue-length prediction key)
};

#define CK(call)                                                                                                   \
    do {                                                                                                           \
        cudaError_t e_ = (call);                                                                                   \
        if (e_ != cudaSuccess) {                                                                                   \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                                         \
            return RL_ERR_CUDA;                                                                                    \
        }                                                                                                          \
    } while (0)

// LBVH + triangle records are staged in shared memory only when small: at 92 KB per CTA (576 triangles) the occupancy
// loss made the traversal 22 % slower than reading the tables through L1 (8.87 vs 7.26 ms per 40 M rays); at 23 KB both are equal.
static constexpr size_t kMaxSmemScene = 32 * 1024;
static constexpr uint32_t kMaxIters = 4096; // wavefront iterations per batch (path depth); 0.95^4096 ~ 1e-91
#ifndef RL_SYNC_GROUP
#define RL_SYNC_GROUP 4
#endif

static int grid_for(const rl_ctx *ctx, size_t n, int per_sm, int block = kBlock) {
    size_t blocks = (n + block - 1) / block;
    size_t cap = (size_t)ctx->sm_count * per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

// Grid sizes (measured on B200, cbox 1024^2 x 32 spp, per-stage CUDA events; tools/ab_env.py RL_GRID_PER_SM):
//   group-table traversal kernels: grid-stride over the queue with kTravPerSm CTAs per SM (the tree kernels, which
//   walk per-warp chunks of the queue, use 16).  Rays differ in cost, so MORE CTAs than
//   are resident (5 per SM for k_trace_flat) balance better through the hardware block scheduler: 4/SM 5.01 ms,
//   resident (5) 4.93, 8 4.77, 16 4.68, 32 4.60, 64 4.56 ms -- 32 keeps the per-CTA scene staging (5 KB) negligible.
//   k_shade: one wave of resident CTAs (4 per SM at 64 registers); 6/SM 5.52 ms vs 4.81 ms.
static constexpr int kTravPerSm = 32;
// Resident CTAs per SM of a kernel at a block size / dynamic shared memory size (occupancy API), cached per (kernel, block, smem).
static int resident_per_sm(const void *kernel, size_t smem, int block = kBlock) {
    struct Key {
        const void *k;
        size_t smem;
        int block, nb;
    };
    static std::mutex mu;
    static std::vector<Key> cache;
    std::lock_guard<std::mutex> lock(mu);
    for (auto &c : cache)
        if (c.k == kernel && c.smem == smem && c.block == block) return c.nb;
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, block, smem) != cudaSuccess || nb < 1) nb = 1;
    cache.push_back({kernel, smem, block, nb});
    return nb;
}
static int tree_per_sm() {
    static int v = 0;
    if (!v) {
        v = 16; // tess24 (20 736 triangles): 4 / 8 / 16 / 32 CTAs per SM: 27.7 / 27.3 / 26.7 / 26.8 ms
        if (const char *e = getenv("RL_TREE_PER_SM")) v = std::max(1, atoi(e)); // A/B hook
    }
    return v;
}
// Queue length at which k_tail takes a batch over: one resident wave of its CTAs (4 x 128 threads per SM at 128 registers).
// Measured (tools/tail_ab.sh, cbox 1024^2 x 32 spp / one rank of 8 at 128 spp): off 11.20 / 5.96 ms, 37 888 10.94, 75 776 10.95 / 5.67,
// 151 552 10.97 / 5.74, 303 104 11.21 / 5.91 ms -- beyond one wave the 128-register kernel is no faster than the wavefront.
static size_t tail_threshold(const rl_ctx *ctx) {
    if (const char *e = getenv("RL_TAIL_MAX")) return (size_t)std::max(0L, atol(e)); // A/B + test hook, read per render; 0 = pure wavefront
    return (size_t)ctx->sm_count * 512;
}
static int trav_per_sm() {
    static int v = 0;
    if (!v) {
        v = kTravPerSm;
        if (const char *e = getenv("RL_GRID_PER_SM")) v = std::max(1, atoi(e)); // A/B hook
    }
    return v;
}

// The definitions below take C linkage from their declarations in rl_b200.h.

int rl_abi_version(void) { return RL_B200_ABI_VERSION; }

const char *rl_last_error(const rl_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int rl_nccl_unique_id(void *out_128_bytes) {
    if (!out_128_bytes) return RL_ERR_INVALID;
    if (!load_nccl(g_create_error)) return RL_ERR_NCCL;
    int rc = g_nccl.GetUniqueId(out_128_bytes);
    if (rc != 0) {
        g_create_error = std::string("ncclGetUniqueId: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error");
        return RL_ERR_NCCL;
    }
    return RL_OK;
}

int rl_create(rl_ctx **out, int device, int nranks, int rank, const void *nccl_unique_id) {
    if (!out || nranks < 1 || rank < 0 || rank >= nranks) {
        g_create_error = "rl_create: bad arguments";
        return RL_ERR_INVALID;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(e) + " (this library has no CPU path)";
        return RL_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) {
        g_create_error = "rl_create: device ordinal out of range";
        return RL_ERR_INVALID;
    }
    rl_ctx *ctx = new rl_ctx;
    ctx->device = device, ctx->nranks = nranks, ctx->rank = rank;
    auto fail = [&](int code) { // everything created so far goes through the same teardown as a live context
        g_create_error = ctx->err;
        rl_destroy(ctx);
        return code;
    };
#define CKC(call)                                                                                                  \
    do {                                                                                                           \
        cudaError_t e_ = (call);                                                                                   \
        if (e_ != cudaSuccess) {                                                                                   \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                                         \
            return fail(RL_ERR_CUDA);                                                                              \
        }                                                                                                          \
    } while (0)
    CKC(cudaSetDevice(device));
    cudaDeviceProp prop;
    CKC(cudaGetDeviceProperties(&prop, device));
    ctx->sm_count = prop.multiProcessorCount;
    ctx->mem_total = prop.totalGlobalMem;
    CKC(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    { // keep up to 512 MB of freed scene memory in the stream-ordered pool instead of returning it to the driver at every synchronisation
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t thr = 512ull << 20;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
        }
    }
    CKC(cudaMalloc(&ctx->d_counts, 4 * sizeof(uint32_t)));
    CKC(cudaMalloc(&ctx->d_counters, sizeof(Counters)));
    CKC(cudaMalloc(&ctx->d_hist, 4 * kMaxIters * sizeof(uint32_t))); // queue lengths, shadow-queue lengths, fix-list lengths (closest, shadow) per iteration
    CKC(cudaMallocHost(&ctx->h_hist, 2 * kMaxIters * sizeof(uint32_t)));
    CKC(cudaMallocHost(&ctx->h_counts, 4 * sizeof(uint32_t)));
    CKC(cudaMallocHost(&ctx->h_counters, sizeof(Counters)));
    for (auto &ev : ctx->ev) CKC(cudaEventCreate(&ev));
    CKC(cudaFuncSetAttribute(k_trace<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmemScene));
    CKC(cudaFuncSetAttribute(k_shadow<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmemScene));
#undef CKC
    if (nranks > 1 && nccl_unique_id) {
        if (!load_nccl(ctx->err)) return fail(RL_ERR_NCCL);
        NcclId id;
        std::memcpy(id.internal, nccl_unique_id, 128);
        // ncclCommInitRank(ncclComm_t*, int nranks, ncclUniqueId commId /*by value*/, int rank)
        typedef int (*init_fn)(ncclComm_t *, int, NcclId, int);
        int rc = ((init_fn)g_nccl.CommInitRank)(&ctx->comm, nranks, id, rank);
        if (rc != 0) {
            ctx->err = std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error");
            return fail(RL_ERR_NCCL);
        }
    }
    *out = ctx;
    return RL_OK;
}

void *rl_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
void rl_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

void rl_destroy(rl_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->comm);
    for (int i = 0; i < 2; i++) {
        cudaFree(ctx->ray_o[i]);
        cudaFree(ctx->ray_d[i]);
        cudaFree(ctx->state[i]);
    }
    cudaFree(ctx->hit), cudaFree(ctx->sh_a), cudaFree(ctx->sh_b), cudaFree(ctx->sh_c), cudaFree(ctx->lacc);
    cudaFree(ctx->fix_trace), cudaFree(ctx->fix_shadow);
    cudaFree(ctx->img_sum), cudaFree(ctx->pixel_list), cudaFree(ctx->frame);
    cudaFree(ctx->d_counts), cudaFree(ctx->d_counters), cudaFree(ctx->d_hist);
    cudaFreeHost(ctx->h_hist);
    cudaFreeHost(ctx->h_counts), cudaFreeHost(ctx->h_counters);
    for (auto &ev : ctx->ev)
        if (ev) cudaEventDestroy(ev);
    for (auto &ev : ctx->ev_pool) cudaEventDestroy(ev);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int rl_set_profiling(rl_ctx *ctx, int on) {
    if (!ctx) return RL_ERR_INVALID;
    ctx->profiling = on < 0 ? 0 : (on > 2 ? 2 : on);
    return RL_OK;
}

int rl_layout(rl_ctx *ctx, rl_layout_info *out) {
    if (!ctx || !out) return RL_ERR_INVALID;
    out->ray_bytes = 32, out->hit_bytes = 16, out->state_bytes = 16, out->shadow_bytes = 48, out->accum_bytes = 16;
    out->max_paths_in_flight = ctx->cap_paths;
    return RL_OK;
}

// ---- scene -------------------------------------------------------------------------------------------
template <typename T>
static cudaError_t upload(T **dst, const std::vector<T> &src, cudaStream_t st) {
    cudaError_t e = cudaMallocAsync(dst, std::max<size_t>(src.size(), 1) * sizeof(T), st); // stream-ordered pool: no device-wide synchronisation (e2e rebuilds the scene every step)
    if (e != cudaSuccess) return e;
    if (!src.empty()) e = cudaMemcpyAsync(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice, st);
    return e;
}

// scene memory comes from the device's stream-ordered pool (cudaMallocAsync on the context's stream; release threshold raised in rl_create)
#define SFREE(p) ((p) == nullptr ? cudaSuccess : (ctx ? cudaFreeAsync((p), ctx->stream) : cudaFree(p)))
void rl_scene_destroy(rl_ctx *ctx, rl_scene *s) {
    if (!s) return;
    if (ctx) cudaSetDevice(ctx->device);
    SFREE(s->d_trav), SFREE(s->d_nodes), SFREE(s->d_shade), SFREE(s->d_verts), SFREE(s->d_mats);
    SFREE(s->d_emit_info), SFREE(s->d_emit_cdf), SFREE(s->d_area_cdf), SFREE(s->d_flat);
    SFREE(s->d_uvs), SFREE(s->d_tex), SFREE(s->d_texels), SFREE(s->d_env_dist), SFREE(s->d_ats_nodes), SFREE(s->d_ats_leaf);
    SFREE(s->d_quad_verts), SFREE(s->d_cam_masks);
    SFREE(s->d_ref_nodes), SFREE(s->d_ref_prims), SFREE(s->d_ref_up);
    delete s;
}

int rl_scene_create(rl_ctx *ctx, const rl_scene_desc *desc, rl_scene **out) {
    if (!ctx || !desc || !out) return RL_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    rl_scene *s = new rl_scene;
    static uint32_t scene_counter = 0;
    s->scene_gen = ++scene_counter;
    std::string err;
    if (!build_host_scene(desc, s->hs, err)) {
        ctx->err = "rl_scene_create: " + err;
        bool unsupported = desc && (desc->has_volume || desc->has_environment > 1);
        delete s;
        return unsupported ? RL_ERR_UNSUPPORTED : RL_ERR_INVALID;
    }
    HostScene &hs = s->hs;
    const uint32_t n = hs.ntris;
    cudaStream_t st = ctx->stream;
    uint64_t *d_keys = nullptr, *d_keys_sorted = nullptr;
    int2 *d_children = nullptr;
    int2v *d_ranges = nullptr;
    int *d_parent_node = nullptr, *d_parent_leaf = nullptr, *d_flags = nullptr;
    float4 *d_leaf_lo = nullptr, *d_leaf_hi = nullptr, *d_node_lo = nullptr, *d_node_hi = nullptr;
    void *d_tmp = nullptr;
    auto cleanup = [&]() {
        SFREE(d_keys), SFREE(d_keys_sorted), SFREE(d_children), SFREE(d_ranges), SFREE(d_parent_node), SFREE(d_parent_leaf);
        SFREE(d_flags), SFREE(d_leaf_lo), SFREE(d_leaf_hi), SFREE(d_node_lo), SFREE(d_node_hi), SFREE(d_tmp);
    };
#define CKS(call)                                                                                                  \
    do {                                                                                                           \
        cudaError_t e_ = (call);                                                                                   \
        if (e_ != cudaSuccess) {                                                                                   \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                                         \
            cleanup();                                                                                             \
            rl_scene_destroy(ctx, s);                                                                              \
            return RL_ERR_CUDA;                                                                                    \
        }                                                                                                          \
    } while (0)
    CKS(upload(&s->d_verts, hs.verts, st));
    CKS(upload(&s->d_shade, hs.shade, st));
    CKS(upload(&s->d_mats, hs.mats, st));
    CKS(upload(&s->d_emit_info, hs.emit_info, st));
    CKS(upload(&s->d_emit_cdf, hs.emit_cdf, st));
    CKS(upload(&s->d_area_cdf, hs.area_cdf, st));
    if (!hs.uvs.empty()) CKS(upload(&s->d_uvs, hs.uvs, st));
    if (!hs.tex.empty()) CKS(upload(&s->d_tex, hs.tex, st));
    if (!hs.texels.empty()) CKS(upload(&s->d_texels, hs.texels, st));
    if (!hs.env_dist.empty()) CKS(upload(&s->d_env_dist, hs.env_dist, st));
    if (!hs.ats_nodes.empty()) {
        CKS(upload(&s->d_ats_nodes, hs.ats_nodes, st));
        CKS(upload(&s->d_ats_leaf, hs.ats_leaf_of_prim, st));
    }
    const uint32_t n_nodes = n > 1 ? n - 1 : 1;
    CKS(cudaMallocAsync(&s->d_trav, (size_t)n * RL_TRAV_F4 * sizeof(float4), st));
    CKS(cudaMallocAsync(&d_keys, (size_t)n * 8, st));
    CKS(cudaMallocAsync(&d_keys_sorted, (size_t)n * 8, st));
    CKS(cudaMallocAsync(&d_leaf_lo, (size_t)n * sizeof(float4), st));
    CKS(cudaMallocAsync(&d_leaf_hi, (size_t)n * sizeof(float4), st));
    // The tree collapses subtrees of <= 2 triangles into leaves.  Scenes of <= 64 triangles additionally get the group
    // table (rl_flat_host.hpp), which every ray scans instead of walking the tree.
    int leaf_max = 2; // measured on a 20 736-triangle scene (tools/tess_cbox.py 24): leaves of <= 1 / 2 / 4 / 8 / 16 triangles: 31.2 / 27.4 / 28.4 / 31.1 / 36.1 ms
    if (const char *e = getenv("RL_LEAF_MAX")) leaf_max = std::max(1, std::min(RL_LEAF_MAX_CAP, atoi(e)));
    bool flat_ok = n <= (uint32_t)RL_LEAF_MAX_CAP;
    if (const char *e = getenv("RL_FLAT")) flat_ok = flat_ok && atoi(e) != 0;
    if (n >= (1u << 25)) {
        ctx->err = "rl_scene_create: more than 2^25 triangles";
        cleanup();
        rl_scene_destroy(ctx, s);
        return RL_ERR_UNSUPPORTED;
    }
    // The reference's own tree (BVHAccel::new: full SAH sweep, leaves of <= 2 primitives), rebuilt on the host.  It decides ties and rim
    // hits (rl_refbvh_host.hpp) and -- for scenes without a group table -- its TOPOLOGY is also the tree the traversal kernels walk:
    // a sweep-SAH tree costs incoherent rays far fewer node visits than the Morton-order LBVH (measured below), and the build is host work
    // the reference pays too.  The device boxes stay conservative (the reference's boxes grown by bvh_box_eps), so results do not change.
    RefBVH rb;
    bool have_rb = false;
    if (getenv("RL_NO_REF_ORDER") == nullptr) { // (A/B + test hook: lowest-index tie rule of NaiveAcceleration instead)
        build_ref_bvh(hs, rb);
        have_rb = rb.depth + 2 <= (uint32_t)RL_STACK_SIZE;
    }
    bool use_sah = have_rb && n > (uint32_t)RL_LEAF_MAX_CAP && n > (uint32_t)leaf_max;
    if (const char *e = getenv("RL_TREE")) use_sah = have_rb && n > 2u && n > (uint32_t)leaf_max && std::strcmp(e, "lbvh") != 0; // A/B: lbvh | sah
    if (use_sah) leaf_max = std::max(leaf_max, 2);
    // Morton keys over the raw vertex bounds
    V3 smin = V3{hs.raw_min[0], hs.raw_min[1], hs.raw_min[2]};
    V3 ext = V3{hs.raw_max[0] - hs.raw_min[0], hs.raw_max[1] - hs.raw_min[1], hs.raw_max[2] - hs.raw_min[2]};
    V3 sinv = V3{ext.x > 0.0f ? 1.0f / ext.x : 0.0f, ext.y > 0.0f ? 1.0f / ext.y : 0.0f, ext.z > 0.0f ? 1.0f / ext.z : 0.0f};
    std::vector<uint64_t> h_keys; // slot order of the triangles (Morton order, or the leaf order of the reference's tree)
    if (use_sah) {
        h_keys.resize(n);
        for (uint32_t i = 0; i < n; i++) h_keys[i] = ((uint64_t)i << 32) | (uint64_t)rb.prims[i];
        CKS(cudaMemcpyAsync(d_keys_sorted, h_keys.data(), (size_t)n * 8, cudaMemcpyHostToDevice, st));
    } else {
        k_morton<<<grid_for(ctx, n, 8), kBlock, 0, st>>>(s->d_verts, n, smin, sinv, d_keys);
        CKS(cudaGetLastError());
        size_t tmp_bytes = 0;
        CKS(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, d_keys, d_keys_sorted, (int)n, 0, 64, st));
        CKS(cudaMallocAsync(&d_tmp, std::max<size_t>(tmp_bytes, 16), st));
        CKS(cub::DeviceRadixSort::SortKeys(d_tmp, tmp_bytes, d_keys, d_keys_sorted, (int)n, 0, 64, st));
    }
    k_tri_setup<<<grid_for(ctx, n, 8), kBlock, 0, st>>>(s->d_verts, d_keys_sorted, n, bvh_box_eps(hs.abs_max), s->d_trav, s->d_shade, d_leaf_lo, d_leaf_hi);
    CKS(cudaGetLastError());
    std::vector<int2> h_children;
    std::vector<int2v> h_ranges;
    WideTree wt;
    if (use_sah) { // wide nodes from the reference's topology (rl_wide_host.hpp)
        uint32_t width = 4;
        if (const char *e = getenv("RL_TREE")) width = std::strcmp(e, "sah2") == 0 ? 2u : 4u; // A/B: lbvh | sah2 | sah (= 4-wide)
        build_wide_tree(rb, bvh_box_eps(hs.abs_max), width, wt);
        if (wt.max_stack > (uint32_t)RL_STACK_SIZE) build_wide_tree(rb, bvh_box_eps(hs.abs_max), 2u, wt); // (rb.depth + 2 <= RL_STACK_SIZE was checked)
        CKS(cudaMallocAsync(&s->d_nodes, wt.nodes.size() * sizeof(float4), st));
        CKS(cudaMemcpyAsync(s->d_nodes, wt.nodes.data(), wt.nodes.size() * sizeof(float4), cudaMemcpyHostToDevice, st));
    } else {
        CKS(cudaMallocAsync(&s->d_nodes, (size_t)n_nodes * 4 * sizeof(float4), st));
        h_keys.resize(n); // Morton order of the triangles: group table, and the slot of every primitive for the reference-order tree
        CKS(cudaMemcpyAsync(h_keys.data(), d_keys_sorted, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    }
    if (n > 1 && !use_sah) {
        CKS(cudaMallocAsync(&d_children, (size_t)(n - 1) * sizeof(int2), st));
        CKS(cudaMallocAsync(&d_ranges, (size_t)(n - 1) * sizeof(int2v), st));
        CKS(cudaMallocAsync(&d_parent_node, (size_t)(n - 1) * sizeof(int), st));
        CKS(cudaMallocAsync(&d_parent_leaf, (size_t)n * sizeof(int), st));
        CKS(cudaMallocAsync(&d_flags, (size_t)(n - 1) * sizeof(int), st));
        CKS(cudaMallocAsync(&d_node_lo, (size_t)(n - 1) * sizeof(float4), st));
        CKS(cudaMallocAsync(&d_node_hi, (size_t)(n - 1) * sizeof(float4), st));
        CKS(cudaMemsetAsync(d_flags, 0, (size_t)(n - 1) * sizeof(int), st));
        k_karras<<<grid_for(ctx, n - 1, 8), kBlock, 0, st>>>(d_keys_sorted, (int)n, d_children, d_ranges, d_parent_node, d_parent_leaf);
        CKS(cudaGetLastError());
        k_fit<<<grid_for(ctx, n, 8), kBlock, 0, st>>>((int)n, leaf_max, d_children, d_ranges, d_parent_node, d_parent_leaf, d_leaf_lo, d_leaf_hi, d_node_lo,
                                                        d_node_hi, d_flags, s->d_nodes);
        CKS(cudaGetLastError());
        h_children.resize(n - 1);
        h_ranges.resize(n - 1);
        CKS(cudaMemcpyAsync(h_children.data(), d_children, (size_t)(n - 1) * sizeof(int2), cudaMemcpyDeviceToHost, st));
        CKS(cudaMemcpyAsync(h_ranges.data(), d_ranges, (size_t)(n - 1) * sizeof(int2v), cudaMemcpyDeviceToHost, st));
    }
    CKS(cudaStreamSynchronize(st));
    std::vector<uint32_t> prim_of_slot(n);
    for (uint32_t i = 0; i < n; i++) prim_of_slot[i] = (uint32_t)(h_keys[i] & 0xffffffffull);
    bool ref_ok = false;
    {
        if (have_rb) {
            std::vector<uint32_t> slot_of_prim(n);
            for (uint32_t i = 0; i < n; i++) slot_of_prim[prim_of_slot[i]] = i;
            for (auto &p : rb.prims) p = slot_of_prim[p];
            { // leaf of every triangle, indexed by Morton slot
                const size_t nn = rb.nodes.size() / 2;
                std::vector<uint32_t> leaf_of_slot(n);
                for (uint32_t i = 0; i < n; i++) leaf_of_slot[rb.prims[i]] = rb.up[nn + i];
                for (uint32_t i = 0; i < n; i++) rb.up[nn + i] = leaf_of_slot[i];
                s->sv.ref_n_nodes = (uint32_t)nn;
            }
            CKS(upload(&s->d_ref_up, rb.up, st));
            CKS(upload(&s->d_ref_nodes, rb.nodes, st));
            CKS(upload(&s->d_ref_prims, rb.prims, st));
            s->n_ref_f4 = (uint32_t)rb.nodes.size();
            CKS(cudaStreamSynchronize(st)); // rb goes out of scope
            ref_ok = true;
        }
    }
    if (flat_ok) {
        flat_ok = build_flat_table(hs, prim_of_slot, s->flat);
        if (flat_ok) {
            CKS(upload(&s->d_flat, s->flat.f4, st));
            CKS(upload(&s->d_quad_verts, s->flat.quad_verts, st));
            CKS(cudaStreamSynchronize(st));
        }
    }
    cleanup();
#undef CKS
    // tree statistics over the live part of the tree (and the traversal-stack bound)
    uint32_t max_depth = 1, live_nodes = 0, live_leaves = 0;
    int root_ref;
    if (use_sah) {
        root_ref = 0;
        live_nodes = wt.n_nodes, live_leaves = wt.n_leaves, max_depth = wt.depth;
    } else if (n <= (uint32_t)leaf_max) {
        root_ref = leaf_ref(0u, n);
        live_leaves = 1;
    } else {
        root_ref = 0;
        std::vector<std::pair<int, uint32_t>> stack{{0, 1u}};
        while (!stack.empty()) {
            auto [node, depth] = stack.back();
            stack.pop_back();
            live_nodes++;
            max_depth = std::max(max_depth, depth);
            int2 ch = h_children[node];
            for (int c : {ch.x, ch.y}) {
                if (c >= 0 && h_ranges[c].y - h_ranges[c].x + 1 > leaf_max) stack.push_back({c, depth + 1});
                else live_leaves++;
            }
        }
    }
    if (max_depth + 1 > RL_STACK_SIZE) {
        ctx->err = "rl_scene_create: LBVH deeper than the traversal stack";
        rl_scene_destroy(ctx, s);
        return RL_ERR_UNSUPPORTED;
    }
    s->n_node_f4 = use_sah ? (uint32_t)wt.nodes.size() : n_nodes * 4;
    s->n_trav_f4 = n * RL_TRAV_F4;
    s->smem_bytes = (size_t)(s->n_node_f4 + s->n_trav_f4) * sizeof(float4);
    size_t smem_limit = kMaxSmemScene;
    if (const char *e = getenv("RL_SMEM_MAX_KB")) smem_limit = std::min<size_t>(kMaxSmemScene, (size_t)atoi(e) * 1024); // A/B hook
    s->smem_ok = s->smem_bytes <= smem_limit;
    SceneView &sv = s->sv;
    sv.trav = s->d_trav, sv.nodes = s->d_nodes, sv.shade = s->d_shade, sv.verts = s->d_verts, sv.mats = s->d_mats;
    sv.emit_info = s->d_emit_info, sv.emit_cdf = s->d_emit_cdf, sv.area_cdf = s->d_area_cdf;
    sv.uvs = s->d_uvs, sv.tex = s->d_tex, sv.texels = s->d_texels;
    sv.emit_var = hs.emit_var ? 1u : 0u;
    s->emit_var = hs.emit_var;
    sv.ref_nodes = ref_ok ? s->d_ref_nodes : nullptr, sv.ref_prims = ref_ok ? s->d_ref_prims : nullptr, sv.ref_up = ref_ok ? s->d_ref_up : nullptr;
    sv.ntris = n, sv.n_emitters = hs.n_emitters;
    sv.root_ref = root_ref;
    sv.wide4 = use_sah && wt.width == 4u ? 1u : 0u;
    s->root_tree = root_ref;
    s->root_flat = root_ref;
    s->flat_ok = flat_ok;
    s->smem_flat_bytes = (size_t)(s->flat.n_groups * RL_FLAT_F4 + RL_FLAT_TAIL_F4 + s->n_trav_f4) * sizeof(float4);
    sv.flat = s->d_flat, sv.n_groups = 0; // n_groups is set per launch (launch_trace / launch_shadow)
    sv.flat_valid_a = s->flat.valid_a, sv.flat_valid_b = s->flat.valid_b, sv.flat_delta = s->flat.delta;
    if (const char *e = getenv("RL_FLAT_LEAF")) // A/B: the pre-group-table flat path (whole scene as one leaf of single records)
        if (atoi(e) != 0 && n <= (uint32_t)RL_LEAF_MAX_CAP) s->root_flat = leaf_ref(0u, n), s->flat_ok = false;
    // With a group table every ray scans it (measured on B200, cbox 1024^2 x 32 spp: 13.6 ms vs 14.6 ms with camera rays
    // and first shadow rays on the tree); without one, everything walks the tree.  RL_COHERENT_TREE=1|2 restores the split.
    s->coherent_tree = 0;
    if (const char *e = getenv("RL_COHERENT_TREE")) s->coherent_tree = atoi(e);
    sv.root_min = V3{hs.root_min[0], hs.root_min[1], hs.root_min[2]};
    sv.root_max = V3{hs.root_max[0], hs.root_max[1], hs.root_max[2]};
    sv.abs_max = hs.abs_max;
    sv.env_on = hs.env_on ? 1u : 0u, sv.env_color = Col{hs.env_color[0], hs.env_color[1], hs.env_color[2]};
    sv.bs_center = V3{hs.bs_center[0], hs.bs_center[1], hs.bs_center[2]}, sv.bs_radius = hs.bs_radius, sv.env_pdf_sel = hs.env_pdf_sel;
    sv.env_w = hs.env_w, sv.env_h = hs.env_h, sv.env_texel_off = hs.env_texel_off, sv.env_dist = s->d_env_dist, sv.env_func_int = hs.env_func_int;
    sv.ats_nodes = s->d_ats_nodes, sv.ats_leaf_of_prim = s->d_ats_leaf, sv.ats_root = hs.ats_root;
    std::memcpy(sv.s2c, hs.s2c, 64);
    std::memcpy(sv.c2w, hs.c2w, 64);
    sv.cam_pos = V3{hs.cam_pos[0], hs.cam_pos[1], hs.cam_pos[2]};
    sv.img_w = (float)hs.img_w, sv.img_h = (float)hs.img_h;
    s->info.ntris = n, s->info.nnodes = live_nodes, s->info.nleaves = live_leaves, s->info.max_depth = max_depth;
    std::memcpy(s->info.root_min, hs.root_min, 12);
    std::memcpy(s->info.root_max, hs.root_max, 12);
    {
        uint32_t seen = 0;
        for (uint32_t mi = 0; mi < desc->nmeshes; mi++) seen |= 1u << (desc->meshes[mi].mat.kind & 31u);
        s->n_bsdf_kinds = (uint32_t)__builtin_popcount(seen);
        s->kind_mask = seen;
    }
    s->info.smem_resident = s->smem_ok ? 1u : 0u;
    s->info.flat_groups = s->flat_ok ? s->flat.n_groups : 0u, s->info.flat_pairs = s->flat.n_pairs, s->info.flat_singles = s->flat.n_singles;
    s->info.flat_delta = s->flat.delta;
    *out = s;
    return RL_OK;
}

int rl_scene_bvh_info(rl_ctx *ctx, const rl_scene *scene, rl_bvh_info *out) {
    if (!ctx || !scene || !out) return RL_ERR_INVALID;
    *out = scene->info;
    return RL_OK;
}

// ---- buffers -----------------------------------------------------------------------------------------
static int ensure_paths(rl_ctx *ctx, size_t n_ray, size_t n_shadow = 0, size_t n_lacc = 0) {
    if (n_shadow == 0) n_shadow = n_ray;
    if (n_lacc == 0) n_lacc = n_ray;
    if (n_ray > ctx->cap_paths) {
        for (int i = 0; i < 2; i++) {
            cudaFree(ctx->ray_o[i]), cudaFree(ctx->ray_d[i]), cudaFree(ctx->state[i]);
            ctx->ray_o[i] = ctx->ray_d[i] = ctx->state[i] = nullptr;
        }
        cudaFree(ctx->hit), cudaFree(ctx->fix_trace);
        ctx->hit = nullptr, ctx->fix_trace = nullptr;
        ctx->cap_paths = 0;
        size_t bytes = n_ray * sizeof(float4);
        for (int i = 0; i < 2; i++) {
            CK(cudaMalloc(&ctx->ray_o[i], bytes));
            CK(cudaMalloc(&ctx->ray_d[i], bytes));
            CK(cudaMalloc(&ctx->state[i], bytes));
        }
        CK(cudaMalloc(&ctx->hit, bytes));
        CK(cudaMalloc(&ctx->fix_trace, n_ray * sizeof(uint32_t)));
        ctx->cap_paths = n_ray;
    }
    if (n_shadow > ctx->cap_shadow) {
        cudaFree(ctx->sh_a), cudaFree(ctx->sh_b), cudaFree(ctx->sh_c), cudaFree(ctx->fix_shadow);
        ctx->sh_a = ctx->sh_b = ctx->sh_c = nullptr, ctx->fix_shadow = nullptr;
        ctx->cap_shadow = 0;
        CK(cudaMalloc(&ctx->sh_a, n_shadow * sizeof(float4)));
        CK(cudaMalloc(&ctx->sh_b, n_shadow * sizeof(float4)));
        CK(cudaMalloc(&ctx->sh_c, n_shadow * sizeof(float4)));
        CK(cudaMalloc(&ctx->fix_shadow, n_shadow * sizeof(uint32_t)));
        ctx->cap_shadow = n_shadow;
    }
    if (n_lacc > ctx->cap_lacc) {
        cudaFree(ctx->lacc);
        ctx->lacc = nullptr;
        ctx->cap_lacc = 0;
        CK(cudaMalloc(&ctx->lacc, n_lacc * sizeof(float4)));
        ctx->cap_lacc = n_lacc;
    }
    return RL_OK;
}

// Pixels owned by this rank, in 16x16 tile order (tile (tx,ty) -> rank (tx+ty) % nranks).
static int ensure_pixels(rl_ctx *ctx, uint32_t w, uint32_t h) {
    if (ctx->pl_w == w && ctx->pl_h == h && ctx->pixel_list) return RL_OK;
    std::vector<uint32_t> list;
    uint32_t tiles_x = (w + 15) / 16, tiles_y = (h + 15) / 16;
    for (uint32_t ty = 0; ty < tiles_y; ty++)
        for (uint32_t tx = 0; tx < tiles_x; tx++) {
            if (ctx->nranks > 1 && (int)((tx + ty) % (uint32_t)ctx->nranks) != ctx->rank) continue;
            for (uint32_t y = ty * 16; y < std::min(h, ty * 16 + 16); y++)
                for (uint32_t x = tx * 16; x < std::min(w, tx * 16 + 16); x++) list.push_back(y * w + x);
        }
    size_t frame = (size_t)w * h * 3;
    if (list.size() > ctx->cap_pix) {
        cudaFree(ctx->img_sum), cudaFree(ctx->pixel_list);
        ctx->img_sum = nullptr, ctx->pixel_list = nullptr, ctx->cap_pix = 0;
        CK(cudaMalloc(&ctx->img_sum, std::max<size_t>(list.size(), 1) * sizeof(float4)));
        CK(cudaMalloc(&ctx->pixel_list, std::max<size_t>(list.size(), 1) * sizeof(uint32_t)));
        ctx->cap_pix = list.size();
    }
    if (frame > ctx->cap_frame) {
        cudaFree(ctx->frame);
        ctx->frame = nullptr, ctx->cap_frame = 0;
        CK(cudaMalloc(&ctx->frame, (frame + 1) * sizeof(float))); // + 1: the "a rank failed" flag that travels with the reduce
        ctx->cap_frame = frame;
    }
    if (!list.empty()) CK(cudaMemcpy(ctx->pixel_list, list.data(), list.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    ctx->pl_w = w, ctx->pl_h = h, ctx->pl_npix = (uint32_t)list.size();
    ctx->pl_gen++;
    return RL_OK;
}

static int validate(rl_ctx *ctx, const rl_scene *scene, const rl_integrator_desc *I, const rl_render_opts *o) {
    if (!scene || !I || !o) {
        ctx->err = "rl_render: null argument";
        return RL_ERR_INVALID;
    }
    if (o->struct_size != sizeof(rl_render_opts)) {
        ctx->err = "rl_render: rl_render_opts.struct_size mismatch";
        return RL_ERR_INVALID;
    }
    if (o->spp == 0) { // assert_ne!(scene.nb_samples, 0), integrators/mod.rs:410
        ctx->err = "rl_render: nb_samples must not be 0";
        return RL_ERR_INVALID;
    }
    if (o->sampler_mode != RL_SAMPLER_COUNTER) {
        ctx->err = "rl_render: only RL_SAMPLER_COUNTER runs on the GPU (mode A is the oracle's)";
        return RL_ERR_UNSUPPORTED;
    }
    if (I->kind == RL_INTEGRATOR_PATH) {
        if (I->strategy > RL_STRATEGY_EMITTER) {
            ctx->err = "rl_render: bad strategy";
            return RL_ERR_INVALID;
        }
        if (I->max_depth >= 0 && I->max_depth < 2) { // evaluate() unwraps a missing sensor edge: path.rs:154
            ctx->err = "rl_render: max_depth < 2 panics in the reference (path.rs:154)";
            return RL_ERR_INVALID;
        }
        if (scene->hs.n_emitters == 0 && I->strategy != RL_STRATEGY_BSDF) { // scene.rs:97-100
            ctx->err = "rl_render: no emitter in the scene but light sampling requested";
            return RL_ERR_INVALID;
        }
    } else if (I->kind == RL_INTEGRATOR_DIRECT) {
        if (scene->hs.n_emitters == 0 && I->nb_light_samples > 0) {
            ctx->err = "rl_render: no emitter in the scene but light samples requested";
            return RL_ERR_INVALID;
        }
        if (I->nb_bsdf_samples > 64 || I->nb_light_samples > 64) {
            ctx->err = "rl_render: more than 64 bsdf/light samples per pixel sample";
            return RL_ERR_INVALID;
        }
    } else if (I->kind == RL_INTEGRATOR_AO) {
        if (I->ao_max_distance != I->ao_max_distance) {
            ctx->err = "rl_render: ao max_distance is NaN";
            return RL_ERR_INVALID;
        }
    } else {
        ctx->err = "rl_render: unknown integrator kind";
        return RL_ERR_INVALID;
    }
    return RL_OK;
}

// k_fix_flat over the lists a group-table kernel of iteration my_k has just filled (fixc_trace / fixc_shadow: their lengths; nullptr = none)
static void launch_fix(rl_ctx *ctx, rl_scene *sc, const uint32_t *fixc_trace, const uint32_t *fixc_shadow, const float4 *ro, const float4 *rd, float4 *hit,
                       bool camera_origin, const uint32_t *done_at, uint32_t my_k) {
    if (!sc->sv.ref_nodes) return;
    const uint32_t *zero = ctx->d_hist + 2 * kMaxIters - 1; // never written
    SceneView sv = sc->sv;
    const size_t smem = (size_t)(sc->n_ref_f4 + sc->n_trav_f4) * sizeof(float4) + (size_t)sc->hs.ntris * sizeof(uint32_t);
    k_fix_flat<<<ctx->sm_count * 8, 128, smem, ctx->stream>>>(sv, fixc_trace ? fixc_trace : zero, ctx->fix_trace, ro, rd, hit, camera_origin ? 1u : 0u,
                                                          fixc_shadow ? fixc_shadow : zero, ctx->fix_shadow, ctx->sh_a, ctx->sh_b, ctx->sh_c, ctx->lacc, ctx->d_counters,
                                                          done_at, my_k, sc->n_ref_f4, sc->n_trav_f4);
    ctx->launches++;
}
// fix-list lengths of iteration k live behind the queue-length history
static uint32_t *fixc_trace_of(rl_ctx *ctx, uint32_t k) { return ctx->d_hist + 2 * kMaxIters + k; }
static uint32_t *fixc_shadow_of(rl_ctx *ctx, uint32_t k) { return ctx->d_hist + 3 * kMaxIters + k; }

template <bool SMEM>
static void launch_trace(rl_ctx *ctx, rl_scene *sc, const uint32_t *count, size_t n, const float4 *ro, const float4 *rd, float4 *hit,
                         const uint32_t *done_at, uint32_t my_k, bool coherent = false, bool camera_origin = false, const uint32_t *cam_masks = nullptr) {
    SceneView sv = sc->sv;
    const bool tree = coherent && sc->coherent_tree >= 1;
    if (!tree && sc->flat_ok) {
        sv.n_groups = sc->flat.n_groups;
        k_trace_flat<<<grid_for(ctx, n, trav_per_sm()), kBlock, sc->smem_flat_bytes, ctx->stream>>>(sv, count, ro, rd, hit, sc->n_trav_f4, camera_origin ? 1u : 0u, cam_masks,
                                                                                         ctx->pl_npix, done_at, my_k, fixc_trace_of(ctx, my_k), ctx->fix_trace);
        ctx->launches++;
        launch_fix(ctx, sc, fixc_trace_of(ctx, my_k), nullptr, ro, rd, hit, camera_origin, done_at, my_k);
        return;
    }
    sv.root_ref = tree ? sc->root_tree : sc->root_flat;
    k_trace<SMEM><<<grid_for(ctx, n, tree_per_sm()), kBlock, SMEM ? sc->smem_bytes : 0, ctx->stream>>>(sv, count, ro, rd, hit, sc->n_node_f4, sc->n_trav_f4, done_at, my_k);
    ctx->launches++;
}
template <bool SMEM>
static void launch_shadow(rl_ctx *ctx, rl_scene *sc, const uint32_t *count, size_t n, const uint32_t *done_at, uint32_t my_k, bool coherent = false) {
    SceneView sv = sc->sv;
    const bool tree = coherent && sc->coherent_tree >= 2;
    if (!tree && sc->flat_ok) {
        sv.n_groups = sc->flat.n_groups;
        k_shadow_flat<<<grid_for(ctx, n, trav_per_sm()), kBlock, sc->smem_flat_bytes, ctx->stream>>>(sv, count, ctx->sh_a, ctx->sh_b, ctx->sh_c, ctx->lacc, ctx->d_counters,
                                                                                          sc->n_trav_f4, done_at, my_k, fixc_shadow_of(ctx, my_k), ctx->fix_shadow);
        ctx->launches++;
        launch_fix(ctx, sc, nullptr, fixc_shadow_of(ctx, my_k), nullptr, nullptr, nullptr, false, done_at, my_k);
        return;
    }
    sv.root_ref = tree ? sc->root_tree : sc->root_flat;
    k_shadow<SMEM><<<grid_for(ctx, n, tree_per_sm()), kBlock, SMEM ? sc->smem_bytes : 0, ctx->stream>>>(sv, count, ctx->sh_a, ctx->sh_b, ctx->sh_c, ctx->lacc,
                                                                                             ctx->d_counters, sc->n_node_f4, sc->n_trav_f4, done_at, my_k);
    ctx->launches++;
}

// Camera-ray culling masks of a group-table scene for the current pixel list (k_camera_cull), cached per scene.
static int ensure_cam_masks(rl_ctx *ctx, rl_scene *sc) {
    if (sc->cam_gen == ctx->pl_gen && sc->d_cam_masks) return RL_OK;
    const uint32_t npix = ctx->pl_npix, nblk = (npix + 31u) / 32u;
    if (nblk > sc->cam_cap) {
        if (sc->d_cam_masks) cudaFreeAsync(sc->d_cam_masks, ctx->stream);
        sc->d_cam_masks = nullptr, sc->cam_cap = 0;
        CK(cudaMallocAsync(&sc->d_cam_masks, (size_t)std::max(nblk, 1u) * sizeof(uint32_t), ctx->stream));
        sc->cam_cap = nblk;
    }
    const HostScene &hs = sc->hs;
    const double ext = (double)hs.abs_max + std::max(std::max(std::fabs((double)hs.cam_pos[0]), std::fabs((double)hs.cam_pos[1])), std::fabs((double)hs.cam_pos[2]));
    SceneView sv = sc->sv;
    if (npix) {
        k_camera_cull<<<grid_for(ctx, nblk, 8), kBlock, 0, ctx->stream>>>(sv, sc->d_quad_verts, sc->flat.valid_a, ctx->pixel_list, npix, hs.img_w, 1e-4 * ext, sc->d_cam_masks);
        ctx->launches++;
        CK(cudaGetLastError());
    }
    sc->cam_gen = ctx->pl_gen;
    return RL_OK;
}

// ---- per-launch timing without synchronisation ---------------------------------------------------------
// One CUDA event after every launch (and one at the start); the time between two consecutive events is the duration of
// the kernel launched in between (plus its launch gap), attributed to that kernel's stage.  Nothing is read until the
// frame is complete, so the timed kernels are exactly the ones of an untimed frame.
enum EvKind : uint8_t { EV_START = 0, EV_RAYGEN, EV_TRACE, EV_SHADE, EV_SHADOW, EV_TAIL, EV_ACCUM, EV_OTHER };
static void ev_reset(rl_ctx *ctx) { ctx->ev_used = 0; }
static void ev_mark(rl_ctx *ctx, EvKind kind) {
    if (!ctx->profiling) return;
    if (ctx->ev_used == ctx->ev_pool.size()) {
        cudaEvent_t e = nullptr;
        if (cudaEventCreate(&e) != cudaSuccess) return;
        ctx->ev_pool.push_back(e);
        ctx->ev_kind.push_back(EV_OTHER);
    }
    cudaEventRecord(ctx->ev_pool[ctx->ev_used], ctx->stream);
    ctx->ev_kind[ctx->ev_used] = (uint8_t)kind;
    ctx->ev_used++;
}
static void ev_collect(rl_ctx *ctx, rl_stats &S) {
    for (size_t i = 1; i < ctx->ev_used; i++) {
        float ms = 0.0f;
        if (cudaEventElapsedTime(&ms, ctx->ev_pool[i - 1], ctx->ev_pool[i]) != cudaSuccess) continue;
        switch (ctx->ev_kind[i]) {
        case EV_RAYGEN: S.ms_raygen += ms; break;
        case EV_TRACE: S.ms_trace += ms, S.launches_trace++; break;
        case EV_SHADE: S.ms_shade += ms, S.launches_shade++; break;
        case EV_SHADOW: S.ms_shadow += ms; break;
        case EV_TAIL: S.ms_tail += ms; break;
        case EV_ACCUM: S.ms_accum += ms; break;
        default: break;
        }
    }
}

// Queue length expected at wavefront iteration k of a batch of n_paths paths: the previous frame's lengths for the same scene and
// integrator when there is one (averaging wrappers render the same thing again and again), else a geometric model.  Only grid
// sizes and the position of the k_tail window depend on it -- never a result.
static size_t predict_len(const rl_ctx *ctx, size_t n_paths, uint32_t k) {
    double r;
    if (k < ctx->pred_ratio.size()) r = ctx->pred_ratio[k];
    else if (!ctx->pred_ratio.empty()) r = ctx->pred_ratio.back() * std::pow(0.6, (double)(k + 1 - ctx->pred_ratio.size()));
    else r = std::pow(0.62, (double)k);
    return (size_t)std::min<double>((double)n_paths, std::ceil(r * (double)n_paths));
}

static int render_impl(rl_ctx *ctx, rl_scene *sc, const rl_integrator_desc *I, const rl_render_opts *o, rl_stats *stats) {
    int rc = validate(ctx, sc, I, o);
    if (rc != RL_OK) return rc;
    CK(cudaSetDevice(ctx->device));
    const uint32_t W = sc->hs.img_w, H = sc->hs.img_h;
    rc = ensure_pixels(ctx, W, H);
    if (rc != RL_OK) return rc;
    const uint32_t npix = ctx->pl_npix;
    cudaStream_t st = ctx->stream;
    rl_stats S{};
    ctx->launches = 0;
    ev_reset(ctx);
    if (ctx->nranks > 1) CK(cudaMemsetAsync(ctx->frame, 0, (size_t)W * H * 3 * sizeof(float), st)); // one rank: k_finish writes every pixel
    CK(cudaMemsetAsync(ctx->d_counters, 0, sizeof(Counters), st));
    CK(cudaEventRecord(ctx->ev[0], st));
    ev_mark(ctx, EV_START);
    uint32_t hist_iters = 0;       // iterations scheduled in the last batch (entries of d_hist worth reading back)
    size_t hist_paths = 0;
    if (npix > 0) {
        const bool ao = I->kind == RL_INTEGRATOR_AO; // runs on the two stages of `direct`: no light samples, one extension ray
        const bool direct = I->kind == RL_INTEGRATOR_DIRECT || ao;
        const uint32_t nl = ao ? 0u : (direct ? I->nb_light_samples : 1u), nbs = ao ? 1u : (direct ? I->nb_bsdf_samples : 1u);
        const uint32_t n_slots = direct ? 1u + nl + nbs : 1u;
        const uint32_t widest = std::max(std::max(nl, nbs), n_slots);
        uint32_t batch = o->batch_spp;
        if (batch == 0) {
            // Auto: as many samples per pixel in flight as a quarter of the device memory holds (11 float4 queues = 176 B per
            // path; 180 GB HBM3e -> 2^28 paths).  Every batch ends in a tail of ~25 nearly empty wavefront iterations that
            // costs ~1 ms whatever the batch size, so fewer, larger batches are faster (measured, cbox 1024^2 x 128 spp:
            // 16 spp/batch 49.5 ms, 32 45.9, 64 44.0, 128 43.0 ms per frame).
            size_t budget = std::min<size_t>((size_t)1 << 28, std::max<size_t>((size_t)1 << 22, ctx->mem_total / 4 / 176));
            batch = (uint32_t)std::max<size_t>(1, budget / ((size_t)npix * widest));
        }
        batch = std::min(batch, o->spp);
        size_t batch_paths = (size_t)npix * batch;
        for (;;) { // an automatic batch size shrinks when the device memory is shared with something else
            rc = ensure_paths(ctx, batch_paths * std::max(1u, nbs), batch_paths * std::max(1u, nl), batch_paths * n_slots);
            if (rc == RL_OK || o->batch_spp != 0 || batch == 1) break;
            cudaGetLastError(); // clear the allocation error
            batch = (batch + 1) / 2;
            batch_paths = (size_t)npix * batch;
        }
        if (rc != RL_OK) return rc;
        IntegParams ip{};
        ip.kind = I->kind, ip.min_depth = I->min_depth, ip.max_depth = I->max_depth, ip.rr_depth = I->rr_depth;
        ip.strategy = I->strategy, ip.single_scattering = I->single_scattering;
        ip.nb_bsdf_samples = ao ? 1u : I->nb_bsdf_samples, ip.nb_light_samples = ao ? 0u : I->nb_light_samples;
        ip.ao_max_distance = I->ao_max_distance, ip.ao_normal_correction = I->ao_normal_correction;
        ip.seed_h = seed_hash(o->seed);
        ip.npix = npix, ip.img_w = W;
        ip.npix_mul = 0u, ip.npix_shift = 0u;
        if (npix > 1u) { // division by the invariant npix as multiply-high + shift, exact for ids < 2^31 (checked per batch below)
            uint32_t l = 0;
            while (((uint64_t)1 << l) < npix) l++;
            const uint64_t m = (((uint64_t)1 << (31 + l)) / npix) + 1u;
            if (m <= 0xffffffffull && l >= 1) ip.npix_mul = (uint32_t)m, ip.npix_shift = l - 1u;
        }
        const int prof = ctx->profiling; // 0 off, 1 per-stage kernels (trace and shadow launched separately), 2 the kernels of an untimed frame
        const bool sort_on = o->material_sort == 1u || (o->material_sort >= 2u && sc->n_bsdf_kinds > 1u);
        const bool extra = sc->d_tex != nullptr || sc->d_ats_nodes != nullptr || sc->emit_var; // textures (incl. an environment texture), the light tree or uv-dependent emission: the kernels with KM bit 8
        // the prediction of the queue lengths belongs to (scene, integrator): forget it when either changes
        const uint64_t pred_key = (uint64_t)(uintptr_t)sc ^ ((uint64_t)I->kind << 56) ^ ((uint64_t)(uint32_t)I->max_depth << 40) ^ ((uint64_t)(uint32_t)I->rr_depth << 24) ^
                                  ((uint64_t)I->strategy << 20) ^ ((uint64_t)sc->scene_gen << 4);
        if (ctx->pred_key != pred_key) ctx->pred_ratio.clear(), ctx->pred_key = pred_key;
        uint32_t *done_at = ctx->d_counts + 3;
        for (uint32_t s0 = 0; s0 < o->spp; s0 += batch) {
            const uint32_t nb = std::min(batch, o->spp - s0);
            const size_t n_paths = (size_t)npix * nb;
            ip.sample_base = o->sample_offset + s0;
            if (n_paths >= ((size_t)1 << 31)) ip.npix_mul = 0u; // the multiply-high form is exact below 2^31 only
            // group-table scenes: the camera rays' origin (sv.cam_pos), path id (= queue index) and path state are constants
            // that the first k_trace_flat / k_shade fill in themselves: raygen writes 32 B per path (direction, accumulator) instead of 80
            const bool camera_o = sc->flat_ok && sc->coherent_tree == 0;
            const uint32_t *cam_masks = nullptr; // camera rays scan only the quads their block of 32 pixels can see
            if (camera_o && getenv("RL_NO_CAM_CULL") == nullptr) { // (A/B hook)
                rc = ensure_cam_masks(ctx, sc);
                if (rc != RL_OK) return rc;
                cam_masks = sc->d_cam_masks;
            }
            k_raygen<<<grid_for(ctx, n_paths, 8), kBlock, 0, st>>>(sc->sv, ip, ctx->pixel_list, (uint32_t)n_paths, camera_o ? nullptr : ctx->ray_o[0],
                                                                   ctx->ray_d[0], ctx->lacc, n_slots);
            ctx->launches++;
            ev_mark(ctx, EV_RAYGEN);
            uint32_t init[4] = {(uint32_t)n_paths, 0, 0, 0xffffffffu}; // [3] = done_at
            CK(cudaMemcpyAsync(ctx->d_counts, init, sizeof(init), cudaMemcpyHostToDevice, st));
            if (direct) {
                CK(cudaMemsetAsync(ctx->d_hist + 2 * kMaxIters, 0, 2 * kMaxIters * sizeof(uint32_t), st)); // fix-list lengths
                // primary rays -> stage 1 (emission, light samples, BSDF samples) -> shadow rays -> extension rays -> stage 2; nothing is read
                // back: the second stage's grid is sized from its upper bound (every primary ray spawns at most nbs extension rays)
                uint32_t *c_in = ctx->d_counts, *c_out = ctx->d_counts + 1, *c_sh = ctx->d_counts + 2;
                const size_t n = n_paths;
                if (sc->smem_ok) launch_trace<true>(ctx, sc, c_in, n, ctx->ray_o[0], ctx->ray_d[0], ctx->hit, done_at, 0u, true, camera_o, cam_masks);
                else launch_trace<false>(ctx, sc, c_in, n, ctx->ray_o[0], ctx->ray_d[0], ctx->hit, done_at, 0u, true, camera_o, cam_masks);
                ev_mark(ctx, EV_TRACE);
                // light-tree scenes: the first vertex' shading normal travels with every extension ray (EmitterSampler::direct_pdf(.., Some(&its.n_s), ..),
                // direct.rs:158-165) in the stage-1 state queue, which camera rays do not use (their state is a constant)
                float4 *ns_buf = (sc->d_ats_nodes && nbs > 0 && !ao) ? ctx->state[0] : nullptr;
#define RL_LAUNCH_DIRECT1(KM) k_shade_direct1<KM><<<grid_for(ctx, n, 4), kBlock, 0, st>>>(sc->sv, ip, ctx->pixel_list, c_in, (uint32_t)n_paths, ctx->ray_o[0], ctx->ray_d[0], \
        ctx->state[0], ctx->hit, ctx->ray_o[1], ctx->ray_d[1], ctx->state[1], c_out, ctx->sh_a, \
        ctx->sh_b, ctx->sh_c, c_sh, ctx->lacc, ctx->d_counters, camera_o ? 3u : 1u, ns_buf)
                if (sc->kind_mask == 0x1u && !extra) RL_LAUNCH_DIRECT1(0x1u);
                else RL_LAUNCH_DIRECT1(RL_KM_ALL);
#undef RL_LAUNCH_DIRECT1

                ctx->launches++;
                ev_mark(ctx, EV_SHADE);
                if (nl > 0) {
                    if (sc->smem_ok) launch_shadow<true>(ctx, sc, c_sh, n * nl, done_at, 0u, true);
                    else launch_shadow<false>(ctx, sc, c_sh, n * nl, done_at, 0u, true);
                    ev_mark(ctx, EV_SHADOW);
                }
                if (nbs > 0) {
                    const size_t n2 = n * nbs;
                    if (sc->smem_ok) launch_trace<true>(ctx, sc, c_out, n2, ctx->ray_o[1], ctx->ray_d[1], ctx->hit, done_at, 1u);
                    else launch_trace<false>(ctx, sc, c_out, n2, ctx->ray_o[1], ctx->ray_d[1], ctx->hit, done_at, 1u);
                    ev_mark(ctx, EV_TRACE);
                    k_shade_direct2<<<grid_for(ctx, n2, 8), kBlock, 0, st>>>(sc->sv, ip, c_out, ctx->ray_o[1], ctx->ray_d[1], ctx->state[1], ctx->hit, ctx->lacc,
                                                                             ctx->d_counters, ns_buf);
                    ctx->launches++;
                    ev_mark(ctx, EV_SHADE);
                }
                k_tally<<<1, 1, 0, st>>>(ctx->d_counts, nullptr, 2u, done_at, 0u, ctx->d_counters); // rays of the two stages: d_counts[0], d_counts[1]
                ctx->launches++;
            } else {
                // `path`.  Queue lengths live in a per-iteration history on the device (iteration k reads hist[k] and appends to
                // hist[k+1]); the whole batch is enqueued without reading anything back (see rl_kernels.cuh: device-decided schedule).
                uint32_t *qc = ctx->d_hist, *shc = ctx->d_hist + kMaxIters;
                CK(cudaMemsetAsync(ctx->d_hist, 0, 4 * kMaxIters * sizeof(uint32_t), st));
                k_set_u32<<<1, 1, 0, st>>>(qc, (uint32_t)n_paths);
                const bool flat = sc->flat_ok && sc->coherent_tree == 0;
                const bool fuse_ok = prof != 1 && flat && getenv("RL_NO_FUSE") == nullptr;
                const size_t tail_max = tail_threshold(ctx); // 0: never hand over (pure wavefront; test / A-B hook)
                // k_tail window: from two iterations before the predicted hand-over a launch after every iteration, the last one unconditional
                uint32_t k_pred = 0;
                while (k_pred < 200u && predict_len(ctx, n_paths, k_pred) > std::max<size_t>(tail_max, 1)) k_pred++;
                // queue indices a k_tail launch is scheduled for: a wide window around the model's guess, a narrow one around last frame's hand-over
                // (queue lengths of two frames of one scene differ by a fraction of a percent; every launch after the hand-over is a dead ~2 us)
                const bool have_hist = !ctx->pred_ratio.empty();
                // (*r03*: with the previous frame's lengths the window is ONE unconditional launch at the iteration that frame handed over at --
                // k_tail strides over whatever the queue holds, so a queue 1 % above the threshold costs nothing, whereas every window
                // iteration runs trace and shadow unfused and the iteration after the hand-over is six dead launches: ~40 us per frame)
                uint32_t win_lo = have_hist ? std::max(k_pred, 1u) : (k_pred > 2u ? k_pred - 2u : 1u), win_hi = have_hist ? win_lo : k_pred + 6u;
                if (I->max_depth >= 0) win_hi = std::min<uint32_t>(win_hi, (uint32_t)I->max_depth), win_lo = std::min(win_lo, win_hi);
                const uint32_t *zero_count = ctx->d_hist + 2 * kMaxIters - 1; // never written
                int cur = 0;
                size_t ub_prev = n_paths;
                uint32_t k = 0;
                bool legacy = tail_max == 0;
                uint32_t legacy_read = 0;
                for (;; k++) {
                    if (k + 2 >= kMaxIters) {
                        ctx->err = "rl_render: a path exceeded 4095 wavefront iterations (no Russian roulette in a closed scene?)";
                        return RL_ERR_UNSUPPORTED;
                    }
                    const size_t n_ub = legacy ? ub_prev : std::min<size_t>(n_paths, predict_len(ctx, n_paths, k) + predict_len(ctx, n_paths, k) / 4 + 1024);
                    const bool fused = fuse_ok && !legacy && k < win_lo; // rays of iteration k + shadow segments of iteration k-1 in one launch
                    if (fused) {
                        SceneView sv = sc->sv;
                        sv.n_groups = sc->flat.n_groups;
                        const int tb = grid_for(ctx, n_ub, trav_per_sm()), sb = k == 0 ? 0 : grid_for(ctx, ub_prev, trav_per_sm());
                        k_trace_shadow_flat<<<tb + sb, kBlock, sc->smem_flat_bytes, st>>>(sv, qc + k, ctx->ray_o[cur], ctx->ray_d[cur], ctx->hit,
                                                                                          k == 0 ? zero_count : shc + k - 1, ctx->sh_a, ctx->sh_b, ctx->sh_c,
                                                                                          ctx->lacc, ctx->d_counters, sc->n_trav_f4, (uint32_t)tb,
                                                                                          (k == 0 && camera_o) ? 1u : 0u, k == 0 ? cam_masks : nullptr, npix, done_at, k,
                                                                                          fixc_trace_of(ctx, k), ctx->fix_trace, fixc_shadow_of(ctx, k == 0 ? 0 : k - 1), ctx->fix_shadow); // shadow segments of iteration k-1
                        ctx->launches++;
                        launch_fix(ctx, sc, fixc_trace_of(ctx, k), k == 0 ? nullptr : fixc_shadow_of(ctx, k - 1), ctx->ray_o[cur], ctx->ray_d[cur], ctx->hit, k == 0 && camera_o, done_at, k);
                    } else if (sc->smem_ok) launch_trace<true>(ctx, sc, qc + k, n_ub, ctx->ray_o[cur], ctx->ray_d[cur], ctx->hit, done_at, k, k == 0, k == 0 && camera_o, k == 0 ? cam_masks : nullptr);
                    else launch_trace<false>(ctx, sc, qc + k, n_ub, ctx->ray_o[cur], ctx->ray_d[cur], ctx->hit, done_at, k, k == 0, k == 0 && camera_o, k == 0 ? cam_masks : nullptr);
                    ev_mark(ctx, EV_TRACE);
#define RL_LAUNCH_SHADE(SORT, KM)                                                                                                                   \
    k_shade<SORT, KM><<<grid_for(ctx, n_ub, resident_per_sm((const void *)k_shade<SORT, KM>, 0, shade_block(KM)), shade_block(KM)), shade_block(KM), 0, st>>>(sc->sv, ip, ctx->pixel_list, qc + k, ctx->ray_o[cur], ctx->ray_d[cur], ctx->state[cur], \
                                                                 ctx->hit, ctx->ray_o[cur ^ 1], ctx->ray_d[cur ^ 1], ctx->state[cur ^ 1], qc + k + 1,     \
                                                                 ctx->sh_a, ctx->sh_b, ctx->sh_c, shc + k, ctx->lacc, ctx->d_counters, k == 0 ? (camera_o ? 3u : 1u) : 0u, done_at, k)
                    // kernel specialised for the BSDF kinds of the scene: {diffuse}, {diffuse, phong}, everything
                    // (textured scenes take the general kernel: bit 8 of the mask)
                    // (one BSDF kind: no sort.  Gathering the ~11 % of rays that missed into whole warps with the tile sort was measured: shade 3.95 -> 4.48 ms per 80 M vertices)
                    if (sc->kind_mask == 0x1u && !extra) RL_LAUNCH_SHADE(false, 0x1u);
                    else if ((sc->kind_mask & ~0x3u) == 0u && !extra) {
                        if (sort_on) RL_LAUNCH_SHADE(true, 0x3u);
                        else RL_LAUNCH_SHADE(false, 0x3u);
                    } else {
                        if (extra) {
                            if (sort_on) RL_LAUNCH_SHADE(true, RL_KM_ALL);
                            else RL_LAUNCH_SHADE(false, RL_KM_ALL);
                        } else {
                            if (sort_on) RL_LAUNCH_SHADE(true, 0xffu);
                            else RL_LAUNCH_SHADE(false, 0xffu);
                        }
                    }
#undef RL_LAUNCH_SHADE
                    ctx->launches++;
                    ev_mark(ctx, EV_SHADE);
                    // shadow segments of iteration k: traced by the next iteration's fused launch, or here
                    const bool next_fused = fuse_ok && !legacy && k + 1 < win_lo;
                    if (!next_fused) {
                        if (sc->smem_ok) launch_shadow<true>(ctx, sc, shc + k, n_ub, done_at, k, k == 0);
                        else launch_shadow<false>(ctx, sc, shc + k, n_ub, done_at, k, k == 0);
                        ev_mark(ctx, EV_SHADOW);
                    }
                    ub_prev = n_ub;
                    cur ^= 1;
                    if (legacy) { // pure wavefront: the host reads the queue lengths every 8 iterations and stops at an empty queue
                        if ((k & 7u) == 7u) {
                            CK(cudaMemcpyAsync(ctx->h_hist + legacy_read, qc + legacy_read, (k + 2 - legacy_read) * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
                            CK(cudaStreamSynchronize(st));
                            legacy_read = k + 1;
                            ub_prev = ctx->h_hist[k + 1];
                            if (ub_prev == 0) break;
                        }
                        continue;
                    }
                    if (k + 1 >= win_lo) { // queue k+1 may go to k_tail (shadow segments of iteration k are resolved: launched above)
                        const bool last = k + 1 >= win_hi;
                        SceneView sv = sc->sv;
                        if (sc->flat_ok) sv.n_groups = sc->flat.n_groups;
                        else sv.root_ref = sc->root_flat;
                        const size_t tn = last ? std::max<size_t>(predict_len(ctx, n_paths, k + 1), tail_max) : tail_max;
                        const int tg = (int)std::min<size_t>((tn + kTailBlock - 1) / kTailBlock, (size_t)ctx->sm_count * 16);
                        const size_t tsm = sc->flat_ok ? sc->smem_flat_bytes : 0;
                        const uint32_t take_max = last ? 0xffffffffu : (uint32_t)tail_max;
#define RL_LAUNCH_TAIL(KM) \
    k_tail<KM><<<tg, kTailBlock, tsm, st>>>(sv, ip, ctx->pixel_list, qc + k + 1, ctx->ray_o[cur], ctx->ray_d[cur], ctx->state[cur], ctx->lacc, ctx->d_counters, sc->n_trav_f4, k + 1, kMaxIters - 1u, done_at, take_max)
                        if (sc->kind_mask == 0x1u && !extra) RL_LAUNCH_TAIL(0x1u);
                        else if ((sc->kind_mask & ~0x3u) == 0u && !extra) RL_LAUNCH_TAIL(0x3u);
                        else if (extra) RL_LAUNCH_TAIL(RL_KM_ALL);
                        else RL_LAUNCH_TAIL(0xffu);
#undef RL_LAUNCH_TAIL
                        ctx->launches++;
                        ev_mark(ctx, EV_TAIL);
                        if (last) break;
                    }
                }
                hist_iters = std::min<uint32_t>(k + 2, kMaxIters - 1);
                hist_paths = n_paths;
                k_tally<<<1, 1, 0, st>>>(qc, shc, hist_iters, done_at, 0u, ctx->d_counters);
                ctx->launches++;
            }
            k_accum<<<grid_for(ctx, npix, 8), kBlock, 0, st>>>(ctx->lacc, npix, nb, n_slots, ctx->img_sum, s0 == 0 ? 1 : 0);
            ctx->launches++;
            ev_mark(ctx, EV_ACCUM);
        }
        k_finish<<<grid_for(ctx, npix, 8), kBlock, 0, st>>>(ctx->img_sum, ctx->pixel_list, npix, 1.0f / (float)o->spp, ctx->frame);
        ctx->launches++;
        ev_mark(ctx, EV_ACCUM);
    }
    CK(cudaMemcpyAsync(ctx->h_counters, ctx->d_counters, sizeof(Counters), cudaMemcpyDeviceToHost, st));
    if (hist_iters) CK(cudaMemcpyAsync(ctx->h_hist, ctx->d_hist, hist_iters * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(ctx->h_counts, ctx->d_counts, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CK(cudaEventRecord(ctx->ev[1], st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    float ms_total = 0;
    CK(cudaEventElapsedTime(&ms_total, ctx->ev[0], ctx->ev[1]));
    S.ms_total = ms_total;
    ev_collect(ctx, S);
    S.samples = (uint64_t)npix * o->spp;
    const Counters &C = *ctx->h_counters;
    if (C.tail_overflow) { // same limit and message as the wavefront loop
        ctx->err = "rl_render: a path exceeded 4095 wavefront iterations (no Russian roulette in a closed scene?)";
        return RL_ERR_UNSUPPORTED;
    }
    if (hist_iters && hist_paths) { // the last batch's queue lengths size the next frame's grids (up to the hand-over: later entries were never written)
        const uint32_t done = ctx->h_counts[3];
        ctx->pred_ratio.clear();
        for (uint32_t k = 0; k < hist_iters && k <= done; k++) {
            if (ctx->h_hist[k] == 0) break;
            ctx->pred_ratio.push_back((double)ctx->h_hist[k] / (double)hist_paths);
        }
    }
    S.hits = C.hits;
    S.segments = C.wave_segments + C.tail_segments;
    S.max_depth_seen = std::max<uint64_t>(C.wave_iters, C.tail_iters);
    S.shadow_rays = C.nee_sampled;
    S.shadow_visible = C.shadow_visible;
    S.shadow_traced = C.shadow_traced + C.tail_shadow;
    S.kernel_launches = ctx->launches;
    if (stats) *stats = S;
    return RL_OK;
}

int rl_render_device(rl_ctx *ctx, rl_scene *scene, const rl_integrator_desc *integrator, const rl_render_opts *opts, float *out_rgb_device,
                     rl_stats *stats) {
    if (!ctx) return RL_ERR_INVALID;
    rl_stats S{};
    int rc = render_impl(ctx, scene, integrator, opts, &S);
    if (rc != RL_OK) return rc;
    if (out_rgb_device) {
        size_t bytes = (size_t)scene->hs.img_w * scene->hs.img_h * 3 * sizeof(float);
        CK(cudaMemcpyAsync(out_rgb_device, ctx->frame, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    if (stats) *stats = S;
    return RL_OK;
}

int rl_render(rl_ctx *ctx, rl_scene *scene, const rl_integrator_desc *integrator, const rl_render_opts *opts, float *out_rgb, rl_stats *stats) {
    if (!ctx) return RL_ERR_INVALID;
    rl_stats S{};
    const int rc = render_impl(ctx, scene, integrator, opts, &S);
    // Argument errors (validate) are the same on every rank: nobody enters the collective.  Anything later (out of memory, a path
    // over the iteration limit on THIS rank's tiles, a CUDA error) must not leave the peers waiting in ncclReduce: the failing rank
    // contributes a zero frame and raises the flag that travels behind the frame; rank 0 reports it.
    const bool collective = ctx->nranks > 1 && ctx->comm && scene && ctx->frame && (rc == RL_OK || rc != RL_ERR_INVALID);
    if (rc != RL_OK && !collective) return rc;
    const size_t count = (size_t)scene->hs.img_w * scene->hs.img_h * 3;
    float ms = 0;
    if (collective) {
        if (count > ctx->cap_frame) return rc != RL_OK ? rc : RL_ERR_CUDA; // (the frame of this size was never allocated: the failure came first)
        const std::string own_err = ctx->err;
        if (rc != RL_OK) {
            cudaGetLastError();
            CK(cudaMemsetAsync(ctx->frame, 0, count * sizeof(float), ctx->stream));
        }
        const float flag = rc != RL_OK ? 1.0f : 0.0f;
        CK(cudaMemcpyAsync(ctx->frame + count, &flag, sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        // one ncclReduce(sum, f32) of the framebuffer: tiles are disjoint, so the sum is exact
        CK(cudaEventRecord(ctx->ev[6], ctx->stream));
        int nrc = g_nccl.Reduce(ctx->frame, ctx->frame, count + 1, /*ncclFloat32*/ 7, /*ncclSum*/ 0, 0, ctx->comm, ctx->stream);
        if (nrc != 0) {
            ctx->err = std::string("ncclReduce: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(nrc) : "error");
            return RL_ERR_NCCL;
        }
        CK(cudaEventRecord(ctx->ev[7], ctx->stream));
        float failed = 0.0f;
        if (ctx->rank == 0) CK(cudaMemcpyAsync(&failed, ctx->frame + count, sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaEventSynchronize(ctx->ev[7]));
        CK(cudaStreamSynchronize(ctx->stream));
        CK(cudaEventElapsedTime(&ms, ctx->ev[6], ctx->ev[7]));
        S.ms_reduce = ms;
        if (rc != RL_OK) {
            ctx->err = own_err;
            return rc;
        }
        if (failed > 0.0f) {
            ctx->err = "rl_render: " + std::to_string((int)failed) + " peer rank(s) failed to render their tiles; the reduced frame is incomplete";
            return RL_ERR_NCCL;
        }
    }
    if (out_rgb && (ctx->rank == 0 || !ctx->comm)) {
        CK(cudaEventRecord(ctx->ev[6], ctx->stream));
        CK(cudaMemcpyAsync(out_rgb, ctx->frame, count * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaEventRecord(ctx->ev[7], ctx->stream));
        CK(cudaEventSynchronize(ctx->ev[7]));
        CK(cudaEventElapsedTime(&ms, ctx->ev[6], ctx->ev[7]));
        S.ms_d2h = ms;
    }
    if (stats) *stats = S;
    return RL_OK;
}

// ---- Acceleration::{trace, visible} batches -----------------------------------------------------------
static int trace_device_rays(rl_ctx *ctx, rl_scene *sc, size_t n, uint32_t *prim, float *tuv, bool coherent = false) {
    uint32_t cnt[4] = {(uint32_t)n, 0u, 0u, 0xffffffffu}; // [3] = done_at: never
    CK(cudaMemcpyAsync(ctx->d_counts, cnt, sizeof(cnt), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_hist + 2 * kMaxIters, 0, 2 * kMaxIters * sizeof(uint32_t), ctx->stream)); // fix-list lengths
    if (sc->smem_ok) launch_trace<true>(ctx, sc, ctx->d_counts, n, ctx->ray_o[0], ctx->ray_d[0], ctx->hit, ctx->d_counts + 3, 0u, coherent);
    else launch_trace<false>(ctx, sc, ctx->d_counts, n, ctx->ray_o[0], ctx->ray_d[0], ctx->hit, ctx->d_counts + 3, 0u, coherent);
    CK(cudaGetLastError());
    std::vector<float4> h(n);
    CK(cudaMemcpyAsync(h.data(), ctx->hit, n * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (size_t i = 0; i < n; i++) {
        uint32_t p = f2u(h[i].w);
        prim[i] = p;
        if (tuv) {
            tuv[3 * i] = p == RL_MISS ? 0.0f : h[i].x;
            tuv[3 * i + 1] = p == RL_MISS ? 0.0f : h[i].y;
            tuv[3 * i + 2] = p == RL_MISS ? 0.0f : h[i].z;
        }
    }
    return RL_OK;
}

namespace {
struct DevBuf { // freed on every path out of the call
    void *p = nullptr;
    ~DevBuf() { cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
    template <typename T>
    T *as() { return static_cast<T *>(p); }
};
} // namespace

int rl_trace(rl_ctx *ctx, rl_scene *sc, size_t n, const float *o, const float *d, uint32_t *prim, float *tuv) {
    if (!ctx || !sc || !o || !d || !prim) return RL_ERR_INVALID;
    if (n == 0) return RL_OK;
    if (n > 0x7fffffffu) return RL_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    int rc = ensure_paths(ctx, n);
    if (rc != RL_OK) return rc;
    DevBuf d_o, d_d;
    CK(d_o.alloc(n * 12));
    CK(d_d.alloc(n * 12));
    CK(cudaMemcpyAsync(d_o.p, o, n * 12, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_d.p, d, n * 12, cudaMemcpyHostToDevice, ctx->stream));
    k_pack_rays<<<grid_for(ctx, n, 8), kBlock, 0, ctx->stream>>>(d_o.as<float>(), d_d.as<float>(), (uint32_t)n, ctx->ray_o[0], ctx->ray_d[0]);
    CK(cudaGetLastError());
    return trace_device_rays(ctx, sc, n, prim, tuv); // synchronises the stream before the buffers go
}

int rl_primary_hits(rl_ctx *ctx, rl_scene *sc, uint32_t *prim, float *tuv) {
    if (!ctx || !sc || !prim) return RL_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    size_t n = (size_t)sc->hs.img_w * sc->hs.img_h;
    int rc = ensure_paths(ctx, n);
    if (rc != RL_OK) return rc;
    k_primary_rays<<<grid_for(ctx, n, 8), kBlock, 0, ctx->stream>>>(sc->sv, sc->hs.img_w, sc->hs.img_h, ctx->ray_o[0], ctx->ray_d[0]);
    return trace_device_rays(ctx, sc, n, prim, tuv, true); // camera rays: the tree root
}

int rl_visible(rl_ctx *ctx, rl_scene *sc, size_t n, const float *p0, const float *p1, uint8_t *out) {
    if (!ctx || !sc || !p0 || !p1 || !out) return RL_ERR_INVALID;
    if (n == 0) return RL_OK;
    if (n > 0x7fffffffu) return RL_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    DevBuf b_a, b_b, b_out;
    CK(b_a.alloc(n * 12));
    CK(b_b.alloc(n * 12));
    CK(b_out.alloc(n));
    float *d_a = b_a.as<float>(), *d_b = b_b.as<float>();
    unsigned char *d_out = b_out.as<unsigned char>();
    CK(cudaMemcpyAsync(d_a, p0, n * 12, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_b, p1, n * 12, cudaMemcpyHostToDevice, ctx->stream));
    SceneView sv = sc->sv;
    if (sc->flat_ok) sv.n_groups = sc->flat.n_groups; // group table read through L1 (no staging in this small-batch kernel)
    k_visible_batch<<<grid_for(ctx, n, 8), kBlock, 0, ctx->stream>>>(sv, d_a, d_b, (uint32_t)n, d_out);
    cudaError_t e = cudaMemcpyAsync(out, d_out, n, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    else cudaStreamSynchronize(ctx->stream); // nothing may still read the buffers when they are freed
    if (e != cudaSuccess) {
        ctx->err = std::string("rl_visible: ") + cudaGetErrorString(e);
        return RL_ERR_CUDA;
    }
    return RL_OK;
}

