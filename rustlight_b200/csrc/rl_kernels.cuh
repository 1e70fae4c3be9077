// rl_kernels.cuh -- the wavefront kernels (sm_100a).
//
// One stage per kernel over SoA queues in HBM, every record a 16-byte float4 so each access
// is one coalesced 128-bit load/store:
//   ray_o[i]  = {o.xyz, path_id}          ray_d[i] = {d.xyz, pdf of the sampled direction}
//   state[i]  = {throughput.rgb, depth<<16 | rng_n}
//   hit[i]    = {t, u, v, prim}
//   shadow queue: sh_a[k] = {p0.xyz, path_id}  sh_b[k] = {p1.xyz, -}  sh_c[k] = {radiance.rgb, -}
//   lacc[path_id] = {radiance.rgb, -}     per-path radiance accumulator (fixed slot)
// Kernels are persistent: the grid is a multiple of the SM count and each CTA walks tiles of
// 256 records; queue lengths are read from device memory.  Survivors and shadow segments are
// stream-compacted with warp ballots + one atomicAdd per CTA tile.  The BVH and the triangle
// records are staged in shared memory when they fit (Cornell box: 4.5 KB).
//
// Stages <- reference:  k_raygen  <- Path::from_sensor + Camera::generate (paths/path.rs:56-73, camera.rs:81-91)
//                       k_trace   <- Acceleration::trace   (accel.rs:292-315)
//                       k_shade   <- DirectionalSamplingStrategy::bounce + LightSamplingStrategy::sample
//                                    + evalute_edge MIS (directional.rs:44-107, emitters.rs:108-175, path.rs:37-111)
//                       k_shadow  <- Acceleration::visible (accel.rs:316-343) + edge contribution
//                       k_accum / k_finish <- Bitmap::accumulate, scale(1/spp) (structure.rs:397-425, mod.rs:431-436)
#pragma once
#include <cuda_runtime.h>

#include "rl_build.cuh"
#include "rl_device.cuh"

namespace rl {

#ifndef RL_BLOCK
#define RL_BLOCK 256 // threads per CTA of every kernel (A/B hook)
#endif
constexpr int kBlock = RL_BLOCK;

struct Counters { // device-side statistics, 64-bit
    unsigned long long hits, nee_sampled, shadow_visible, pad;
    unsigned long long tail_segments, tail_iters; // k_tail: closest-hit calls; deepest iteration reached (absolute)
    unsigned long long tail_overflow, pad2;       // a path reached the iteration limit inside k_tail
    unsigned long long wave_segments, shadow_traced; // k_tally: rays / shadow segments the wavefront kernels traced
    unsigned long long wave_iters, tail_shadow;      // deepest wavefront iteration that held a ray; shadow segments traced inside k_tail
};

// ---- device-decided schedule -------------------------------------------------------------------------
// The host enqueues a whole batch without reading anything back: wavefront iterations k = 0, 1, ... with grids sized from
// a prediction (the kernels grid-stride over the true queue lengths, which live in device memory), and from the iteration
// where the queue is expected to fit one resident wave a k_tail launch after every iteration that takes the batch over
// as soon as the queue is short enough (or unconditionally, for the last one).  `done_at` = the iteration whose queue a
// tail kernel consumed (0xffffffff: none yet): every kernel of iteration k returns at once when *done_at <= k.
__device__ __forceinline__ bool batch_done(const uint32_t *done_at, uint32_t my_k) { return *reinterpret_cast<const volatile uint32_t *>(done_at) <= my_k; }

__device__ __forceinline__ float4 ldg4(const float4 *p) { return __ldg(p); }

// ---- scene staging -------------------------------------------------------------------------------
// Copies the wide nodes and the traversal records to dynamic shared memory (128-bit copies).
__device__ __forceinline__ void stage_scene(const SceneView &sv, float4 *smem, uint32_t n_node_f4, uint32_t n_trav_f4) {
    for (uint32_t i = threadIdx.x; i < n_node_f4; i += blockDim.x) smem[i] = ldg4(sv.nodes + i);
    for (uint32_t i = threadIdx.x; i < n_trav_f4; i += blockDim.x) smem[n_node_f4 + i] = ldg4(sv.trav + i);
    __syncthreads();
}

// ---- ray generation ------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_raygen(SceneView sv, IntegParams ip, const uint32_t *__restrict__ pixel_list, uint32_t n_paths,
                                                   float4 *__restrict__ ray_o, float4 *__restrict__ ray_d, float4 *__restrict__ lacc,
                                                   uint32_t n_slots) {
    for (uint32_t id = blockIdx.x * blockDim.x + threadIdx.x; id < n_paths; id += gridDim.x * blockDim.x) {
        uint32_t s_local, lp;
        ip_split(ip, id, &s_local, &lp);
        uint32_t pixel = __ldg(pixel_list + lp);
        uint32_t px = pixel % ip.img_w, py = pixel / ip.img_w;
        Sampler smp = make_sampler(ip.seed_h, pixel, ip.sample_base + s_local, 0u);
        float jx = smp.next();
        float jy = smp.next();
        V3 o, d;
        camera_generate(sv, (float)px + jx, (float)py + jy, &o, &d);
        if (ray_o) ray_o[id] = make_float4(o.x, o.y, o.z, u2f(id)); // nullptr: every camera ray starts at sv.cam_pos with path_id == queue index
        ray_d[id] = make_float4(d.x, d.y, d.z, 1.0f);
        // No path-state record: every camera ray starts with throughput 1, depth 1 and two draws taken (smp.n == 2); the first
        // k_shade / k_shade_direct1 fill that in themselves instead of moving 16 B per path through HBM twice.
        for (uint32_t sl = 0; sl < n_slots; sl++) lacc[(size_t)sl * n_paths + id] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
}

// ---- closest-hit traversal ------------------------------------------------------------------------
// Persistent while-while traversal with in-warp ray refill (Aila & Laine 2009): every warp owns a
// contiguous chunk of the queue; lanes whose ray has finished pull the next rays of the chunk
// once at least kRefillIdle lanes are idle (measured best on B200: 32 = refill when the whole warp is
// idle; partial refills pay the per-ray setup at low lane occupancy), so that the descend phase and the leaf
// phase (exact triangle test) each run with most lanes active although rays are incoherent.
#ifndef RL_REFILL_IDLE
#define RL_REFILL_IDLE 24
#endif
#ifndef RL_WHILE_WHILE
#define RL_WHILE_WHILE 1
#endif
#ifndef RL_SELF_FIX
#define RL_SELF_FIX 1 // group-table closest-hit kernels re-trace a thread's first ambiguous ray themselves, after the loop; 0 = every one goes to k_fix_flat (A/B hook)
#endif
constexpr int kRefillIdle = RL_REFILL_IDLE;
#ifndef RL_TREE_MINBLOCKS
#define RL_TREE_MINBLOCKS 4 // resident CTAs per SM the tree kernels are compiled for (64 registers; tess24 x 16 spp, 1 / 4 / 5 / 6 / 8: 24.0 / 21.6 / 24.0 / 27.9 / 32.0 ms)
#endif

__device__ __forceinline__ void warp_chunk(uint32_t n, uint32_t *begin, uint32_t *end) {
    const uint32_t warps = gridDim.x * (blockDim.x >> 5);
    const uint32_t w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    uint32_t per = (n + warps - 1) / warps;
    per = (per + 31u) & ~31u;
    uint64_t b = (uint64_t)w * per;
    *begin = b < n ? (uint32_t)b : n;
    *end = (b + per) < n ? (uint32_t)(b + per) : n;
}

template <bool SMEM>
__global__ void __launch_bounds__(kBlock, RL_TREE_MINBLOCKS) k_trace(SceneView sv, const uint32_t *__restrict__ count, const float4 *__restrict__ ray_o,
                                                  const float4 *__restrict__ ray_d, float4 *__restrict__ hit, uint32_t n_node_f4,
                                                  uint32_t n_trav_f4, const uint32_t *__restrict__ done_at, uint32_t my_k) {
    extern __shared__ float4 smem[];
    if (batch_done(done_at, my_k)) return;
    const float4 *nodes = sv.nodes, *trav = sv.trav;
    if (SMEM) {
        stage_scene(sv, smem, n_node_f4, n_trav_f4);
        nodes = smem;
        trav = smem + n_node_f4;
    }
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t cursor, chunk_end;
    warp_chunk(*count, &cursor, &chunk_end);
    Trav tr;
    int stack[RL_STACK_SIZE];
#if RL_TREE_DIST
    float sdist[RL_STACK_SIZE]; // entry distances of the stacked nodes (trav_pop)
#else
    float *sdist = nullptr;
#endif
    tr.cur = RL_TRAV_DONE;
    uint32_t my = RL_MISS; // queue index of the ray this lane is tracing
    for (;;) {
        const unsigned idle = __ballot_sync(0xffffffffu, tr.cur == RL_TRAV_DONE);
        if (cursor < chunk_end && (__popc(idle) >= kRefillIdle)) {
            if (tr.cur == RL_TRAV_DONE) {
                uint32_t i = cursor + __popc(idle & ((1u << lane) - 1u));
                if (i < chunk_end) {
                    float4 ro = ray_o[i], rd = ray_d[i];
                    if (closest_begin(tr, sv, xyz(ro), xyz(rd))) my = i;
                    else hit[i] = make_float4(RL_F32_MAX, 0.0f, 0.0f, u2f(RL_MISS)); // root box missed
                }
            }
            cursor += __popc(idle);
            continue; // re-evaluate: lanes whose new ray missed the root box are still idle
        }
        if (idle == 0xffffffffu) break; // chunk exhausted and every lane finished
#if RL_WHILE_WHILE
        while ((uint32_t)tr.cur < (uint32_t)RL_TRAV_DONE) trav_node_step<false>(tr, stack, nodes, sdist); // descend: inner nodes
        if (tr.cur < 0) trav_leaf_closest(tr, stack, trav, sdist);                                         // one leaf
#else
        if ((uint32_t)tr.cur < (uint32_t)RL_TRAV_DONE) trav_node_step<false>(tr, stack, nodes, sdist);
        else if (tr.cur < 0) trav_leaf_closest(tr, stack, trav, sdist);
#endif
        if (tr.cur == RL_TRAV_DONE && my != RL_MISS) {
            HitRec h = closest_result(tr);
            if (RL_REF_ORDER && sv.ref_nodes && closest_ambiguous(tr, sv)) ref_bvh_closest(sv, trav, tr.o, tr.d, &h.t, &h.u, &h.v, &h.prim); // tied / rim hits: the reference's own traversal
            hit[my] = make_float4(h.t, h.u, h.v, u2f(h.prim));
            my = RL_MISS;
        }
    }
}

// ---- camera rays of group-table scenes: which quads can a block of 32 pixels see? -------------------------------
// One thread per block of 32 consecutive local pixels (the unit a warp of k_trace_flat traces): the pixel-space bounding
// rectangle of the block (jitter included, grown by 0.05 px), the pyramid of its four corner rays through the camera position
// (Camera::generate in double), and per quad of the table the classic conservative frustum test (rl_device.cuh:
// camera_block_mask).  Culling only removes exact tests that must fail: images are unchanged.
__global__ void __launch_bounds__(kBlock) k_camera_cull(SceneView sv, const float4 *__restrict__ quad_verts, uint32_t valid_quads,
                                                        const uint32_t *__restrict__ pixel_list, uint32_t npix, uint32_t img_w, double eps,
                                                        uint32_t *__restrict__ masks) {
    const uint32_t nblk = (npix + 31u) >> 5;
    for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < nblk; b += gridDim.x * blockDim.x) {
        uint32_t x0 = 0xffffffffu, x1 = 0u, y0 = 0xffffffffu, y1 = 0u;
        for (uint32_t k = 32u * b; k < min(32u * b + 32u, npix); k++) {
            const uint32_t pixel = __ldg(pixel_list + k), px = pixel % img_w, py = pixel / img_w;
            x0 = min(x0, px), x1 = max(x1, px), y0 = min(y0, py), y1 = max(y1, py);
        }
        masks[b] = camera_block_mask(sv, quad_verts, valid_quads, x0, x1, y0, y1, eps);
    }
}

// ---- group-table variants (scenes of a few dozen triangles, incoherent rays) ------------------------
// One ray per thread, grid-stride, no traversal state: the reference's root-box test, the lockstep scan of the
// flat group table (rl_device.cuh: flat_scan, packed f32x2 arithmetic, no divergence), then the exact triangle
// test on the few survivors.  The group table and the exact-test records are staged in shared memory.
__device__ __forceinline__ void stage_flat(const SceneView &sv, float4 *smem, uint32_t n_flat_f4, uint32_t n_trav_f4) {
    for (uint32_t i = threadIdx.x; i < n_flat_f4; i += blockDim.x) smem[i] = ldg4(sv.flat + i);
    for (uint32_t i = threadIdx.x; i < n_trav_f4; i += blockDim.x) smem[n_flat_f4 + i] = ldg4(sv.trav + i);
    __syncthreads();
}
// The two loops as device functions over a virtual grid (bid of nblocks), so that one launch can run both (below).
// `cam_masks` (camera rays only, else nullptr): per block of 32 consecutive local pixels the quads that overlap the frustum of
// those pixels (k_camera_cull).  A warp traces 32 consecutive paths = 32 consecutive local pixels of one sample index, so the
// OR over the warp is (nearly always) one block's mask and whole groups of the table are skipped by the whole warp.
template <bool CULL>
__device__ __forceinline__ void trace_flat_body_t(const SceneView &sv, const float4 *flat, const float4 *trav, uint32_t bid, uint32_t nblocks, uint32_t n,
                                                const float4 *__restrict__ ray_o, const float4 *__restrict__ ray_d, float4 *__restrict__ hit, bool camera,
                                                const uint32_t *__restrict__ cam_masks, uint32_t npix, uint32_t *fix_count, uint32_t *fix_list) {
#if RL_SELF_FIX
    uint32_t pend = RL_MISS; // a ray of this thread whose answer the reference's own traversal must give (ties, rim hits: a few in 10^4)
#endif
    for (uint32_t i = bid * blockDim.x + threadIdx.x; i < n; i += nblocks * blockDim.x) {
        const float4 rd = ray_d[i];
        const V3 o = camera ? sv.cam_pos : xyz(ray_o[i]), d = xyz(rd); // camera rays share their origin: it is not stored (k_raygen)
        const V3 inv = V3{1.0f / d.x, 1.0f / d.y, 1.0f / d.z};
        uint32_t quads = 0xffffffffu;
        if (CULL) quads = __reduce_or_sync(__activemask(), __ldg(cam_masks + ((i % npix) >> 5)));
        HitRec h;
        h.t = RL_F32_MAX, h.u = 0.0f, h.v = 0.0f, h.prim = RL_MISS;
        bool needs_ref = false;
        if (aabb_intersect_ref(sv.root_min, sv.root_max, o, inv, RL_EPSILON, RL_F32_MAX)) h = flat_closest<CULL, true>(sv, flat, trav, o, d, quads, &needs_ref);
        hit[i] = make_float4(h.t, h.u, h.v, u2f(h.prim));
#if RL_SELF_FIX
        if (needs_ref) { // the thread re-traces its first such ray itself once its share of the queue is done; a second one goes to k_fix_flat's list
            if (pend == RL_MISS) pend = i;
            else fix_list[atomicAdd(fix_count, 1u)] = i;
        }
#else
        if (needs_ref) fix_list[atomicAdd(fix_count, 1u)] = i; // a few rays in 10^4 (ties, rim hits): re-traced over the reference's tree by k_fix_flat
#endif
    }
#if RL_SELF_FIX
    // *r03*: k_fix_flat's single-lane walks of a short list are pure latency (12-15 us per wavefront iteration: 3 % of one rank's frame on 8
    // GPUs, 10 % of config 1).  Here the walk of the thread's own ray runs while other warps are still scanning and k_fix_flat, still launched
    // for what is left (shadow segments blocked by rim hits only, second rays), finds its lists empty (~4 us).  Only the walk AFTER the loop is
    // free: a call inside the loop, or one that takes the SceneView by reference (shadow segments), costs every ray ~10 % (measured: trace 3.29 ->
    // 3.56-3.64 ms, shadow 2.33 -> 2.74-2.79 ms per 80 M segments).  The reference's tree is read through L1; the triangle records are the staged ones.
    if (pend != RL_MISS) {
        const V3 o = camera ? sv.cam_pos : xyz(ray_o[pend]), d = xyz(ray_d[pend]);
        HitRec h;
        ref_bvh_closest(sv, trav, o, d, &h.t, &h.u, &h.v, &h.prim); // the ray passed the root test (it had a hit)
        hit[pend] = make_float4(h.t, h.u, h.v, u2f(h.prim));
    }
#endif
}
__device__ __forceinline__ void trace_flat_body(const SceneView &sv, const float4 *flat, const float4 *trav, uint32_t bid, uint32_t nblocks, uint32_t n,
                                                const float4 *__restrict__ ray_o, const float4 *__restrict__ ray_d, float4 *__restrict__ hit, bool camera,
                                                const uint32_t *__restrict__ cam_masks, uint32_t npix, uint32_t *fix_count, uint32_t *fix_list) {
    if (cam_masks) trace_flat_body_t<true>(sv, flat, trav, bid, nblocks, n, ray_o, ray_d, hit, camera, cam_masks, npix, fix_count, fix_list);
    else trace_flat_body_t<false>(sv, flat, trav, bid, nblocks, n, ray_o, ray_d, hit, camera, cam_masks, npix, fix_count, fix_list);
}
__device__ __forceinline__ void shadow_flat_body(const SceneView &sv, const float4 *flat, const float4 *trav, uint32_t bid, uint32_t nblocks, uint32_t n,
                                                 const float4 *__restrict__ sh_a, const float4 *__restrict__ sh_b, const float4 *__restrict__ sh_c,
                                                 float4 *__restrict__ lacc, Counters *counters, uint32_t *fix_count, uint32_t *fix_list) {
    uint32_t c_vis = 0;
    for (uint32_t i = bid * blockDim.x + threadIdx.x; i < n; i += nblocks * blockDim.x) {
        const float4 a = sh_a[i], b = sh_b[i];
        V3 d;
        float thr;
        bool needs_ref = false;
        // root test failed => "not visible" (accel.rs:338-340): nothing to add
        const bool vis = visible_setup(sv, xyz(a), xyz(b), &d, &thr) && !flat_any<true>(sv, flat, trav, xyz(a), d, thr, &needs_ref);
        if (needs_ref) fix_list[atomicAdd(fix_count, 1u)] = i; // blocked by rim hits only: k_fix_flat decides (and adds the contribution)
        if (vis) {
            const float4 c = sh_c[i];
            const uint32_t pid = f2u(a.w);
            float4 l = lacc[pid];
            l.x += c.x, l.y += c.y, l.z += c.z;
            lacc[pid] = l;
            c_vis++;
        }
    }
    for (int off = 16; off > 0; off >>= 1) c_vis += __shfl_down_sync(0xffffffffu, c_vis, off);
    if ((threadIdx.x & 31u) == 0 && c_vis) atomicAdd(&counters->shadow_visible, (unsigned long long)c_vis);
}
#ifndef RL_TRAV_MINBLOCKS
#define RL_TRAV_MINBLOCKS 5 // resident CTAs per SM the group-table kernels are compiled for (register cap; A/B hook)
#endif
__global__ void __launch_bounds__(kBlock, RL_TRAV_MINBLOCKS) k_trace_flat(SceneView sv, const uint32_t *__restrict__ count, const float4 *__restrict__ ray_o,
                                                       const float4 *__restrict__ ray_d, float4 *__restrict__ hit, uint32_t n_trav_f4, uint32_t camera,
                                                       const uint32_t *__restrict__ cam_masks, uint32_t npix, const uint32_t *__restrict__ done_at, uint32_t my_k,
                                                       uint32_t *fix_count, uint32_t *fix_list) {
    extern __shared__ float4 smem[];
    if (batch_done(done_at, my_k) || blockIdx.x * blockDim.x >= *count) return; // nothing for this CTA: skip the staging
    const uint32_t n_flat_f4 = sv.n_groups * RL_FLAT_F4 + RL_FLAT_TAIL_F4;
    stage_flat(sv, smem, n_flat_f4, n_trav_f4);
    trace_flat_body(sv, smem, smem + n_flat_f4, blockIdx.x, gridDim.x, *count, ray_o, ray_d, hit, camera != 0u, cam_masks, npix, fix_count, fix_list);
}
__global__ void __launch_bounds__(kBlock, RL_TRAV_MINBLOCKS) k_shadow_flat(SceneView sv, const uint32_t *__restrict__ count, const float4 *__restrict__ sh_a,
                                                        const float4 *__restrict__ sh_b, const float4 *__restrict__ sh_c, float4 *__restrict__ lacc,
                                                        Counters *counters, uint32_t n_trav_f4, const uint32_t *__restrict__ done_at, uint32_t my_k,
                                                        uint32_t *fix_count, uint32_t *fix_list) {
    extern __shared__ float4 smem[];
    if (batch_done(done_at, my_k) || blockIdx.x * blockDim.x >= *count) return;
    const uint32_t n_flat_f4 = sv.n_groups * RL_FLAT_F4 + RL_FLAT_TAIL_F4;
    stage_flat(sv, smem, n_flat_f4, n_trav_f4);
    shadow_flat_body(sv, smem, smem + n_flat_f4, blockIdx.x, gridDim.x, *count, sh_a, sh_b, sh_c, lacc, counters, fix_count, fix_list);
}
// Extension rays of wavefront iteration k+1 and shadow rays of iteration k are independent (the shadow kernel only adds
// to the per-path accumulators, which the traversal does not touch): one launch runs both, CTAs [0, trace_blocks) on the
// ray queue and the rest on the shadow queue.  Every kernel costs ~7 us whatever its queue length (launch + table staging
// + one scan at minimal occupancy), and a frame has ~40 iterations: two launches per iteration instead of three.
__global__ void __launch_bounds__(kBlock, RL_TRAV_MINBLOCKS) k_trace_shadow_flat(SceneView sv, const uint32_t *__restrict__ count, const float4 *__restrict__ ray_o,
                                                              const float4 *__restrict__ ray_d, float4 *__restrict__ hit,
                                                              const uint32_t *__restrict__ sh_count, const float4 *__restrict__ sh_a,
                                                              const float4 *__restrict__ sh_b, const float4 *__restrict__ sh_c, float4 *__restrict__ lacc,
                                                              Counters *counters, uint32_t n_trav_f4, uint32_t trace_blocks, uint32_t camera,
                                                              const uint32_t *__restrict__ cam_masks, uint32_t npix,
                                                              const uint32_t *__restrict__ done_at, uint32_t my_k, uint32_t *fix_count_trace,
                                                              uint32_t *fix_list_trace, uint32_t *fix_count_shadow, uint32_t *fix_list_shadow) {
    extern __shared__ float4 smem[];
    if (batch_done(done_at, my_k)) return; // (fused launches are only scheduled before the first k_tail launch: both halves are live or neither)
    if (blockIdx.x < trace_blocks ? blockIdx.x * blockDim.x >= *count : (blockIdx.x - trace_blocks) * blockDim.x >= *sh_count) return;
    const uint32_t n_flat_f4 = sv.n_groups * RL_FLAT_F4 + RL_FLAT_TAIL_F4;
    stage_flat(sv, smem, n_flat_f4, n_trav_f4);
    if (blockIdx.x < trace_blocks) trace_flat_body(sv, smem, smem + n_flat_f4, blockIdx.x, trace_blocks, *count, ray_o, ray_d, hit, camera != 0u, cam_masks, npix, fix_count_trace, fix_list_trace);
    else shadow_flat_body(sv, smem, smem + n_flat_f4, blockIdx.x - trace_blocks, gridDim.x - trace_blocks, *sh_count, sh_a, sh_b, sh_c, lacc, counters, fix_count_shadow, fix_list_shadow);
}

// ---- the rays the hot kernels could not decide for certain ------------------------------------------------
// A few rays in 10^4: two accepted hits within the tie window, a hit on the rim of its triangle, or one the reference's boxes may
// lose (rl_device.cuh: hit_unsafe).  For them the reference's answer is whatever BVHAccel::intersect does, so they are re-traced
// over the reference's own tree (ref_bvh_closest / ref_bvh_any): closest-hit rays get their hit record rewritten, shadow segments
// are decided here and their contribution added.  Runs right after the kernel that filled the lists, before anything reads the hits.
__global__ void __launch_bounds__(128) k_fix_flat(SceneView sv, const uint32_t *__restrict__ n_trace, const uint32_t *__restrict__ list_trace,
                                                  const float4 *__restrict__ ray_o, const float4 *__restrict__ ray_d, float4 *__restrict__ hit, uint32_t camera,
                                                  const uint32_t *__restrict__ n_shadow, const uint32_t *__restrict__ list_shadow, const float4 *__restrict__ sh_a,
                                                  const float4 *__restrict__ sh_b, const float4 *__restrict__ sh_c, float4 *__restrict__ lacc, Counters *counters,
                                                  const uint32_t *__restrict__ done_at, uint32_t my_k, uint32_t n_ref_f4, uint32_t n_trav_f4) {
    extern __shared__ float4 smem[];
    if (batch_done(done_at, my_k)) return;
    // ONE ray per warp (lane 0): the walk is irregular, 32 different walks in one warp would execute one after the other
    // (measured: 25-50 us per launch whatever the list length, against ~10 us like this)
    const uint32_t nt = *n_trace, ns = *n_shadow;
    const uint32_t warps_per_cta = blockDim.x >> 5, warp0 = blockIdx.x * warps_per_cta, n_warps = gridDim.x * warps_per_cta;
    if (warp0 >= max(nt, ns)) return; // (usually: the lists hold a handful of rays)
    // the reference's tree, its leaf contents and the triangle records in shared memory: the walk is a chain of dependent loads
    float4 *s_nodes = smem, *s_trav = smem + n_ref_f4;
    uint32_t *s_prims = reinterpret_cast<uint32_t *>(smem + n_ref_f4 + n_trav_f4);
    for (uint32_t i = threadIdx.x; i < n_ref_f4; i += blockDim.x) s_nodes[i] = ldg4(sv.ref_nodes + i);
    for (uint32_t i = threadIdx.x; i < n_trav_f4; i += blockDim.x) s_trav[i] = ldg4(sv.trav + i);
    for (uint32_t i = threadIdx.x; i < sv.ntris; i += blockDim.x) s_prims[i] = __ldg(sv.ref_prims + i);
    __syncthreads();
    sv.ref_nodes = s_nodes, sv.ref_prims = s_prims;
    if ((threadIdx.x & 31u) != 0u) return;
    const uint32_t w = warp0 + (threadIdx.x >> 5);
    for (uint32_t j = w; j < nt; j += n_warps) {
        const uint32_t i = list_trace[j];
        const float4 rd = ray_d[i];
        const V3 o = camera ? sv.cam_pos : xyz(ray_o[i]), d = xyz(rd);
        HitRec h;
        ref_bvh_closest(sv, s_trav, o, d, &h.t, &h.u, &h.v, &h.prim); // the ray passed the root test (it had a hit)
        hit[i] = make_float4(h.t, h.u, h.v, u2f(h.prim));
    }
    for (uint32_t j = w; j < ns; j += n_warps) {
        const uint32_t i = list_shadow[j];
        const float4 a = sh_a[i], b = sh_b[i];
        V3 d;
        float thr;
        if (visible_setup(sv, xyz(a), xyz(b), &d, &thr) && !ref_bvh_any(sv, s_trav, xyz(a), d, thr)) {
            const float4 c = sh_c[i];
            const uint32_t pid = f2u(a.w);
            float4 l = lacc[pid];
            l.x += c.x, l.y += c.y, l.z += c.z;
            lacc[pid] = l;
            atomicAdd(&counters->shadow_visible, 1ull);
        }
    }
}

// ---- block-level compaction helper ---------------------------------------------------------------
// Returns this thread's slot in the output queue (valid when flag), reserving the CTA's range
// with one atomicAdd.  `scratch` is 2*(kBlock/32)+2 uint32 of shared memory.
__device__ __forceinline__ uint32_t block_compact(bool flag, uint32_t *global_count, uint32_t *warp_tot, uint32_t *base_slot) {
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    unsigned m = __ballot_sync(0xffffffffu, flag);
    uint32_t in_warp = __popc(m & ((1u << lane) - 1u));
    if (lane == 0) warp_tot[warp] = __popc(m);
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t tot = 0;
        for (int w = 0; w < kBlock / 32; w++) {
            uint32_t c = warp_tot[w];
            warp_tot[w] = tot;
            tot += c;
        }
        *base_slot = tot ? atomicAdd(global_count, tot) : 0u;
    }
    __syncthreads();
    uint32_t slot = *base_slot + warp_tot[warp] + in_warp;
    __syncthreads(); // scratch is reused by the next call
    return slot;
}

// Two queues at once (survivors and shadow segments): one pair of barriers instead of three per queue.
#ifndef RL_COMPACT_2BUF
#define RL_COMPACT_2BUF 1
#endif
template <int B>
__device__ __forceinline__ void block_compact2(bool fa, bool fb, uint32_t *count_a, uint32_t *count_b, uint32_t *scratch /* 2*(B/32)+2 */,
                                               uint32_t *slot_a, uint32_t *slot_b) {
    constexpr int W = B / 32;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const unsigned ma = __ballot_sync(0xffffffffu, fa), mb = __ballot_sync(0xffffffffu, fb);
    const uint32_t lt = (1u << lane) - 1u;
    if (lane == 0) {
        scratch[warp] = __popc(ma);
        scratch[W + warp] = __popc(mb);
    }
    __syncthreads();
    if (threadIdx.x < 2) { // thread 0 scans queue a, thread 1 queue b
        uint32_t *w = scratch + threadIdx.x * W;
        uint32_t tot = 0;
        for (int k = 0; k < W; k++) {
            uint32_t c = w[k];
            w[k] = tot;
            tot += c;
        }
        scratch[2 * W + threadIdx.x] = tot ? atomicAdd(threadIdx.x == 0 ? count_a : count_b, tot) : 0u;
    }
    __syncthreads();
    *slot_a = scratch[2 * W] + scratch[warp] + __popc(ma & lt);
    *slot_b = scratch[2 * W + 1] + scratch[W + warp] + __popc(mb & lt);
#if !RL_COMPACT_2BUF
    // a fast warp could overwrite scratch[warp] with the NEXT tile's count before a slow warp has read this tile's values:
    __syncthreads();
#endif
    // RL_COMPACT_2BUF: the caller alternates between two scratch areas from tile to tile instead.  A warp can write area A again (tile
    // t + 2) only after the second barrier of tile t + 1, which every thread reaches after it has read its slots of tile t.
}

// ---- shade: surface interaction, arrival emission, BSDF sample + RR, NEE sample ------------------
// ---- material sort inside a CTA tile ------------------------------------------------------------
// Counting sort of the tile's 256 records by key (0..5 = rl_bsdf_kind of the surface hit, 6 miss, 7 = past the end
// of the queue) with warp ballots + per-warp prefix sums in shared memory.  Returns the tile-local index of
// the record this thread should process, so that every warp shades one material kind (and rays that
// missed are grouped into whole warps that exit at once).  `perm` is kBlock uint16, `cnt` kSortKeys*(kBlock/32).
constexpr int kSortKeys = 8;
template <int B>
__device__ __forceinline__ uint32_t tile_sort_by_key(uint32_t key, unsigned short *perm, uint32_t *cnt) {
    constexpr int W = B / 32;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, lt = (1u << lane) - 1u;
    const unsigned mine = __match_any_sync(0xffffffffu, key); // lanes of this warp with my key
#pragma unroll
    for (int k = 0; k < kSortKeys; k++) {
        const unsigned mk = __ballot_sync(0xffffffffu, key == (uint32_t)k);
        if (lane == (uint32_t)k) cnt[k * W + warp] = __popc(mk);
    }
    __syncthreads();
    if (threadIdx.x == 0) { // exclusive scan in key-major, warp-minor order (64 entries)
        uint32_t tot = 0;
        for (int q = 0; q < kSortKeys * W; q++) {
            uint32_t c = cnt[q];
            cnt[q] = tot;
            tot += c;
        }
    }
    __syncthreads();
    uint32_t rank = cnt[key * W + warp] + __popc(mine & lt);
    perm[rank] = (unsigned short)threadIdx.x;
    __syncthreads();
    return perm[threadIdx.x];
}

// CTA shape of k_shade (measured, tools/ab_variants.py + ab_mixed.py + config_times.py): the diffuse-only kernel runs
// 128 threads x 8 CTAs per SM (a tile barrier waits for 4 warps instead of 8: cbox shade 4.42 -> 4.27 ms per 80 M
// vertices); kernels that carry Phong / microfacet code stay at 256 x 4 (Phong walls 55.9 vs 57.7 ms, the mixed scene
// 13.9 vs 18.3 ms sorted: larger tiles sort into fuller warps, and the heavier kernels spill into L1).
#ifndef RL_SHADE_BLOCK_DIFFUSE
#define RL_SHADE_BLOCK_DIFFUSE 128
#endif
#ifndef RL_SHADE_MINBLOCKS_DIFFUSE
#define RL_SHADE_MINBLOCKS_DIFFUSE (1024 / RL_SHADE_BLOCK_DIFFUSE)
#endif
__host__ __device__ constexpr int shade_block(uint32_t km) { return km == 0x1u ? RL_SHADE_BLOCK_DIFFUSE : kBlock; }

// Resident CTAs per SM (= the register cap): {diffuse} 8 x 128 threads and {diffuse, phong} 4 x 256 at 64 registers (Phong walls,
// 512^2 x 512 spp: 56.4 ms; 80 registers 59.0, 110 registers 62.8 ms); the kernels with the microfacet / Fresnel code run best
// unconstrained at 2 x 256 (122 registers, no spills: mixed scene sorted 14.0 -> 13.3 ms, unsorted 21.3 -> 20.1 ms).
__host__ __device__ constexpr int shade_minblocks(uint32_t km) { return km == 0x1u ? RL_SHADE_MINBLOCKS_DIFFUSE : (km == 0x3u ? 4 : 2); }
template <bool SORT, uint32_t KM>
__global__ void __launch_bounds__(shade_block(KM), shade_minblocks(KM)) k_shade(SceneView sv, IntegParams ip, const uint32_t *__restrict__ pixel_list,
                                                  const uint32_t *__restrict__ count_in, const float4 *__restrict__ ray_o,
                                                  const float4 *__restrict__ ray_d, const float4 *__restrict__ state,
                                                  const float4 *__restrict__ hit, float4 *__restrict__ out_o, float4 *__restrict__ out_d,
                                                  float4 *__restrict__ out_state, uint32_t *count_out, float4 *__restrict__ sh_a,
                                                  float4 *__restrict__ sh_b, float4 *__restrict__ sh_c, uint32_t *count_shadow,
                                                  float4 *__restrict__ lacc, Counters *counters, uint32_t primary,
                                                  const uint32_t *__restrict__ done_at, uint32_t my_k) {
    constexpr int B = shade_block(KM);
    if (batch_done(done_at, my_k)) return;
    __shared__ uint32_t s_scratch[2 * (2 * (B / 32) + 2)]; // two areas, used alternately (block_compact2)
    __shared__ unsigned short s_perm[SORT ? B : 1];
    __shared__ uint32_t s_cnt[SORT ? kSortKeys * (B / 32) : 1];
    const uint32_t n = *count_in;
    const uint32_t n_tiles = (n + B - 1) / B;
    uint32_t c_hits = 0, c_nee = 0;
    bool flip = false;
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        uint32_t i = tile * B + threadIdx.x;
        if (SORT) {
            uint32_t key = 7u;
            if (i < n) {
                const uint32_t prim = f2u(hit[i].w);
                key = prim == RL_MISS ? 6u : min(f2u(__ldg(&sv.mats[RL_MAT_F4 * f2u(__ldg(&sv.shade[4 * prim]).w)]).w), 5u);
            }
            const uint32_t src = tile_sort_by_key<B>(key, s_perm, s_cnt);
            i = tile * B + src;
        }
        StepOut so;
        so.alive = false;
        so.shadow = false;
        uint32_t pid = 0;
        if (i < n) {
            // camera rays (k_raygen): bit 0 = constant path state, bit 1 = origin sv.cam_pos and path_id == queue index, neither stored
            const float4 rd = ray_d[i], h4 = hit[i];
            const float4 ro = (primary & 2u) ? make_float4(sv.cam_pos.x, sv.cam_pos.y, sv.cam_pos.z, u2f(i)) : ray_o[i];
            const float4 st4 = (primary & 1u) ? make_float4(1.0f, 1.0f, 1.0f, u2f((1u << 16) | 2u)) : state[i];
            HitRec h;
            h.t = h4.x, h.u = h4.y, h.v = h4.z, h.prim = f2u(h4.w);
            PathState st;
            st.T = Col{st4.x, st4.y, st4.z};
            st.pdf_prev = rd.w;
            st.path_id = f2u(ro.w);
            pid = st.path_id;
            uint32_t packed = f2u(st4.w);
            st.depth = packed >> 16;
            st.rng_n = packed & 0xffffu;
            uint32_t s_local, lp;
            ip_split(ip, st.path_id, &s_local, &lp);
            uint32_t pixel = __ldg(pixel_list + lp);
            path_step<KM>(sv, ip, xyz(ro), xyz(rd), h, st, pixel, ip.sample_base + s_local, &so);
            if (h.prim != RL_MISS) c_hits++;
            if (so.nee_sampled) c_nee++;
            if (so.has_add) {
                float4 l = lacc[st.path_id];
                l.x += so.add.r, l.y += so.add.g, l.z += so.add.b;
                lacc[st.path_id] = l;
            }
            if (so.alive && (so.next.depth >= 0xfff0u || so.next.rng_n >= 0xfff0u)) so.alive = false; // packing guard
        }
        uint32_t slot, sslot;
        block_compact2<B>(so.alive, so.shadow, count_out, count_shadow, s_scratch + (flip ? 2 * (B / 32) + 2 : 0), &slot, &sslot);
        flip = !flip;
        if (so.alive) {
            out_o[slot] = make_float4(so.next_o.x, so.next_o.y, so.next_o.z, u2f(so.next.path_id));
            out_d[slot] = make_float4(so.next_d.x, so.next_d.y, so.next_d.z, so.next.pdf_prev);
            out_state[slot] = make_float4(so.next.T.r, so.next.T.g, so.next.T.b, u2f((so.next.depth << 16) | so.next.rng_n));
        }
        if (so.shadow) {
            sh_a[sslot] = make_float4(so.sh_p0.x, so.sh_p0.y, so.sh_p0.z, u2f(pid));
            sh_b[sslot] = make_float4(so.sh_p1.x, so.sh_p1.y, so.sh_p1.z, 0.0f);
            sh_c[sslot] = make_float4(so.sh_contrib.r, so.sh_contrib.g, so.sh_contrib.b, 0.0f);
        }
    }
    // per-CTA statistics: warp reduce, one atomic per warp
    for (int off = 16; off > 0; off >>= 1) {
        c_hits += __shfl_down_sync(0xffffffffu, c_hits, off);
        c_nee += __shfl_down_sync(0xffffffffu, c_nee, off);
    }
    if ((threadIdx.x & 31u) == 0) {
        if (c_hits) atomicAdd(&counters->hits, (unsigned long long)c_hits);
        if (c_nee) atomicAdd(&counters->nee_sampled, (unsigned long long)c_nee);
    }
}

// ---- tail of a batch: one thread follows one path to its end ----------------------------------------
// Once the ray queue fits one resident wave of the device, a wavefront iteration is all fixed cost (two launches at
// ~7 us each whatever the queue length, plus the host's look at the queue length every few iterations), and a batch still
// has ~25 iterations to go before its longest path ends.  k_tail takes the queue over at that point: each thread runs
// trace -> path_step -> shadow test for its path until the path dies, with the path's accumulator in registers.  Same
// device functions, same per-path order of additions (arrival emission of vertex k, light sample of vertex k, ...), same
// random numbers: images are bit-identical to the pure wavefront (tests/test_gpu_render.py::test_tail_kernel_*).
// Shadow segments still queued by the last k_shade must be resolved before this kernel starts (the host launches
// k_shadow_flat / k_shadow first, same stream).
constexpr int kTailBlock = 128;
#ifndef RL_TAIL_MINBLOCKS
#define RL_TAIL_MINBLOCKS 1 // resident CTAs per SM k_tail is compiled for (register cap; A/B hook)
#endif
template <uint32_t KM>
__global__ void __launch_bounds__(kTailBlock, RL_TAIL_MINBLOCKS) k_tail(SceneView sv, IntegParams ip, const uint32_t *__restrict__ pixel_list,
                                                     const uint32_t *__restrict__ count, const float4 *__restrict__ ray_o,
                                                     const float4 *__restrict__ ray_d, const float4 *__restrict__ state, float4 *__restrict__ lacc,
                                                     Counters *counters, uint32_t n_trav_f4, uint32_t iter_base, uint32_t iter_limit,
                                                     uint32_t *done_at, uint32_t take_max) {
    extern __shared__ float4 smem[];
    // iter_base = the iteration whose queue this launch may consume.  It does when no earlier launch has and the queue holds at
    // most take_max rays (0xffffffff: unconditionally -- the last launch of the schedule).
    if (*reinterpret_cast<volatile uint32_t *>(done_at) < iter_base) return;
    if (*count > take_max) return;
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicMin(done_at, iter_base); // CTAs of THIS launch that start later still see done_at == iter_base: not "<"
    if (blockIdx.x * blockDim.x >= *count) return;
    const float4 *nodes = sv.nodes, *trav = sv.trav, *flat = sv.flat;
    if (sv.n_groups) { // group table + exact-test records in shared memory, as in k_trace_flat
        const uint32_t n_flat_f4 = sv.n_groups * RL_FLAT_F4 + RL_FLAT_TAIL_F4;
        stage_flat(sv, smem, n_flat_f4, n_trav_f4);
        flat = smem;
        trav = smem + n_flat_f4;
    }
    const uint32_t n = *count;
    uint32_t c_hits = 0, c_nee = 0, c_vis = 0, c_seg = 0, c_iters = 0, c_over = 0, c_sht = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 ro = ray_o[i], rd = ray_d[i], st4 = state[i];
        V3 o = xyz(ro), d = xyz(rd);
        PathState st;
        st.T = Col{st4.x, st4.y, st4.z};
        st.pdf_prev = rd.w;
        st.path_id = f2u(ro.w);
        st.depth = f2u(st4.w) >> 16;
        st.rng_n = f2u(st4.w) & 0xffffu;
        uint32_t s_local, lp;
        ip_split(ip, st.path_id, &s_local, &lp);
        const uint32_t pixel = __ldg(pixel_list + lp), sample = ip.sample_base + s_local;
        float4 l = lacc[st.path_id];
        for (uint32_t it = 1;; it++) {
            const HitRec h = trace_closest(sv, flat, nodes, trav, o, d);
            c_seg++;
            StepOut so;
            path_step<KM>(sv, ip, o, d, h, st, pixel, sample, &so);
            if (h.prim != RL_MISS) c_hits++;
            if (so.nee_sampled) c_nee++;
            if (so.has_add) l.x += so.add.r, l.y += so.add.g, l.z += so.add.b;
            if (so.shadow) {
                c_sht++;
                if (trace_visible(sv, flat, nodes, trav, so.sh_p0, so.sh_p1)) {
                    l.x += so.sh_contrib.r, l.y += so.sh_contrib.g, l.z += so.sh_contrib.b;
                    c_vis++;
                }
            }
            c_iters = max(c_iters, it);
            if (!so.alive || so.next.depth >= 0xfff0u || so.next.rng_n >= 0xfff0u) break; // packing guard as in k_shade
            if (iter_base + it >= iter_limit) { // the wavefront's iteration limit (rl_render reports it as an error)
                c_over = 1u;
                break;
            }
            o = so.next_o, d = so.next_d;
            st = so.next;
        }
        lacc[st.path_id] = l;
    }
    for (int off = 16; off > 0; off >>= 1) {
        c_hits += __shfl_down_sync(0xffffffffu, c_hits, off);
        c_nee += __shfl_down_sync(0xffffffffu, c_nee, off);
        c_vis += __shfl_down_sync(0xffffffffu, c_vis, off);
        c_seg += __shfl_down_sync(0xffffffffu, c_seg, off);
        c_sht += __shfl_down_sync(0xffffffffu, c_sht, off);
        c_iters = max(c_iters, __shfl_down_sync(0xffffffffu, c_iters, off));
    }
    if ((threadIdx.x & 31u) == 0) {
        if (c_hits) atomicAdd(&counters->hits, (unsigned long long)c_hits);
        if (c_nee) atomicAdd(&counters->nee_sampled, (unsigned long long)c_nee);
        if (c_vis) atomicAdd(&counters->shadow_visible, (unsigned long long)c_vis);
        if (c_seg) atomicAdd(&counters->tail_segments, (unsigned long long)c_seg);
        if (c_sht) atomicAdd(&counters->tail_shadow, (unsigned long long)c_sht);
        if (c_iters) atomicMax(&counters->tail_iters, (unsigned long long)(iter_base + c_iters));
    }
    if (c_over) atomicMax(&counters->tail_overflow, 1ull);
}

// ---- `direct` integrator, stage 1: primary hit -> emission, light samples, BSDF samples ---------------
// 4 CTAs per SM (64 registers) beat the unconstrained 119 registers / 2 CTAs for the general kernel: direct -b 1 -l 1 at 2048^2 x 16 spp
// shade 5.58 -> 4.85 ms, ao 3.72 -> 3.19 ms (tools/ab_direct.py); 3 CTAs (80 registers) were slower than both (6.47 ms).
#ifndef RL_DIRECT1_MINBLOCKS
#define RL_DIRECT1_MINBLOCKS 4
#endif
// KM: the BSDF kinds of the scene as for k_shade ({diffuse} and "everything" are instantiated): the Cornell-box kernel carries no Phong / microfacet /
// blend / texture / light-tree code and fits its 64 registers without spills (stack frame 216 -> 0 B; shade 5.35 -> 3.63 ms per 2048^2 x 16 spp).
template <uint32_t KM>
__global__ void __launch_bounds__(kBlock, RL_DIRECT1_MINBLOCKS) k_shade_direct1(SceneView sv, IntegParams ip, const uint32_t *__restrict__ pixel_list,
                                                          const uint32_t *__restrict__ count_in, uint32_t n_paths, const float4 *__restrict__ ray_o,
                                                          const float4 *__restrict__ ray_d, const float4 *__restrict__ state,
                                                          const float4 *__restrict__ hit, float4 *__restrict__ out_o, float4 *__restrict__ out_d,
                                                          float4 *__restrict__ out_state, uint32_t *count_out, float4 *__restrict__ sh_a,
                                                          float4 *__restrict__ sh_b, float4 *__restrict__ sh_c, uint32_t *count_shadow,
                                                          float4 *__restrict__ lacc, Counters *counters, uint32_t camera, float4 *__restrict__ out_ns) {
    // out_ns (light-tree scenes, else nullptr): the shading normal of the first vertex per extension ray, for the tree's pdf in stage 2
    __shared__ uint32_t s_warp[kBlock / 32];
    __shared__ uint32_t s_base;
    const uint32_t n = *count_in;
    const uint32_t n_tiles = (n + kBlock - 1) / kBlock;
    uint32_t c_hits = 0, c_nee = 0;
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint32_t i = tile * kBlock + threadIdx.x;
        DirectCtx cx;
        cx.ok = false;
        cx.env_primary = false;
        uint32_t pid = 0;
        if (i < n) {
            // camera-ray constants are not stored (k_raygen): bit 0 = two draws taken, bit 1 = origin sv.cam_pos and path_id == queue index
            const float4 rd = ray_d[i], h4 = hit[i];
            const float4 ro = (camera & 2u) ? make_float4(sv.cam_pos.x, sv.cam_pos.y, sv.cam_pos.z, u2f(i)) : ray_o[i];
            const float4 st4 = (camera & 1u) ? make_float4(1.0f, 1.0f, 1.0f, u2f((1u << 16) | 2u)) : state[i];
            HitRec h;
            h.t = h4.x, h.u = h4.y, h.v = h4.z, h.prim = f2u(h4.w);
            pid = f2u(ro.w);
            uint32_t s_local, lp;
            ip_split(ip, pid, &s_local, &lp);
            if (ip.kind == 2u) ao_begin(sv, ip, xyz(ro), xyz(rd), h, f2u(st4.w) & 0xffffu, __ldg(pixel_list + lp), ip.sample_base + s_local, &cx);
            else direct_begin<KM>(sv, ip, xyz(ro), xyz(rd), h, f2u(st4.w) & 0xffffu, __ldg(pixel_list + lp), ip.sample_base + s_local, &cx);
            if (h.prim != RL_MISS) c_hits++;
            if ((cx.ok || cx.env_primary) && !is_zero(cx.emit)) lacc[pid] = make_float4(cx.emit.r, cx.emit.g, cx.emit.b, 0.0f);
        }
        for (uint32_t j = 0; j < ip.nb_light_samples; j++) {
            V3 p1 = V3{0.0f, 0.0f, 0.0f};
            Col c = Col{0.0f, 0.0f, 0.0f};
            bool valid = false;
            bool emit_sh = cx.ok && direct_light_sample<KM>(sv, &cx, &p1, &c, &valid);
            if (valid) c_nee++;
            uint32_t slot = block_compact(emit_sh, count_shadow, s_warp, &s_base);
            if (emit_sh) {
                sh_a[slot] = make_float4(cx.its.p.x, cx.its.p.y, cx.its.p.z, u2f((1u + j) * n_paths + pid));
                sh_b[slot] = make_float4(p1.x, p1.y, p1.z, 0.0f);
                sh_c[slot] = make_float4(c.r, c.g, c.b, 0.0f);
            }
        }
        for (uint32_t k = 0; k < ip.nb_bsdf_samples; k++) {
            V3 dir = V3{0.0f, 0.0f, 0.0f};
            Col w = Col{0.0f, 0.0f, 0.0f};
            float pdf = 0.0f;
            bool go = cx.ok && (ip.kind == 2u ? ao_sample(ip, &cx, &dir) : direct_bsdf_sample<KM>(&cx, &dir, &w, &pdf));
            uint32_t slot = block_compact(go, count_out, s_warp, &s_base);
            if (go) {
                out_o[slot] = make_float4(cx.its.p.x, cx.its.p.y, cx.its.p.z, u2f(pid));
                out_d[slot] = make_float4(dir.x, dir.y, dir.z, pdf);
                out_state[slot] = make_float4(w.r, w.g, w.b, u2f((1u + ip.nb_light_samples + k) * n_paths + pid));
                if (out_ns) out_ns[slot] = make_float4(cx.its.n_s.x, cx.its.n_s.y, cx.its.n_s.z, 0.0f);
            }
        }
    }
    for (int off = 16; off > 0; off >>= 1) {
        c_hits += __shfl_down_sync(0xffffffffu, c_hits, off);
        c_nee += __shfl_down_sync(0xffffffffu, c_nee, off);
    }
    if ((threadIdx.x & 31u) == 0) {
        if (c_hits) atomicAdd(&counters->hits, (unsigned long long)c_hits);
        if (c_nee) atomicAdd(&counters->nee_sampled, (unsigned long long)c_nee);
    }
}
// stage 2: the BSDF-sampled ray hit something; MIS-weighted emission if it is a light (direct.rs:145-181)
__global__ void __launch_bounds__(kBlock) k_shade_direct2(SceneView sv, IntegParams ip, const uint32_t *__restrict__ count_in,
                                                          const float4 *__restrict__ ray_o, const float4 *__restrict__ ray_d,
                                                          const float4 *__restrict__ state, const float4 *__restrict__ hit, float4 *__restrict__ lacc,
                                                          Counters *counters, const float4 *__restrict__ ns_in) {
    const uint32_t n = *count_in;
    uint32_t c_hits = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float4 ro = ray_o[i], rd = ray_d[i], st4 = state[i], h4 = hit[i];
        HitRec h;
        h.t = h4.x, h.u = h4.y, h.v = h4.z, h.prim = f2u(h4.w);
        if (h.prim != RL_MISS) c_hits++;
        Col c;
        const V3 ns = ns_in ? xyz(ns_in[i]) : V3{0.0f, 0.0f, 0.0f};
        const bool add = ip.kind == 2u ? ao_finish(ip, h, &c) : direct_finish(sv, ip, xyz(ro), xyz(rd), h, Col{st4.x, st4.y, st4.z}, rd.w, &c, ns, ns_in != nullptr);
        if (add) lacc[f2u(st4.w)] = make_float4(c.r, c.g, c.b, 0.0f);
    }
    for (int off = 16; off > 0; off >>= 1) c_hits += __shfl_down_sync(0xffffffffu, c_hits, off);
    if ((threadIdx.x & 31u) == 0 && c_hits) atomicAdd(&counters->hits, (unsigned long long)c_hits);
}

// ---- shadow rays + NEE resolve ---------------------------------------------------------------------
template <bool SMEM>
__global__ void __launch_bounds__(kBlock, RL_TREE_MINBLOCKS) k_shadow(SceneView sv, const uint32_t *__restrict__ count, const float4 *__restrict__ sh_a,
                                                   const float4 *__restrict__ sh_b, const float4 *__restrict__ sh_c,
                                                   float4 *__restrict__ lacc, Counters *counters, uint32_t n_node_f4, uint32_t n_trav_f4,
                                                   const uint32_t *__restrict__ done_at, uint32_t my_k) {
    extern __shared__ float4 smem[];
    if (batch_done(done_at, my_k)) return;
    const float4 *nodes = sv.nodes, *trav = sv.trav;
    if (SMEM) {
        stage_scene(sv, smem, n_node_f4, n_trav_f4);
        nodes = smem;
        trav = smem + n_node_f4;
    }
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t cursor, chunk_end;
    warp_chunk(*count, &cursor, &chunk_end);
    Trav tr;
    int stack[RL_STACK_SIZE];
    tr.cur = RL_TRAV_DONE;
    uint32_t my = RL_MISS;
    bool blocked = false;
    uint32_t c_vis = 0;
    for (;;) {
        const unsigned idle = __ballot_sync(0xffffffffu, tr.cur == RL_TRAV_DONE);
        if (cursor < chunk_end && (__popc(idle) >= kRefillIdle)) {
            if (tr.cur == RL_TRAV_DONE) {
                uint32_t i = cursor + __popc(idle & ((1u << lane) - 1u));
                if (i < chunk_end) {
                    float4 a = sh_a[i], b = sh_b[i];
                    bool decided, vis;
                    visible_begin(tr, sv, xyz(a), xyz(b), &decided, &vis);
                    blocked = false;
                    if (!decided) my = i; // else: the root test says "not visible" (accel.rs:338-340): nothing to add
                }
            }
            cursor += __popc(idle);
            continue;
        }
        if (idle == 0xffffffffu) break;
#if RL_WHILE_WHILE
        while ((uint32_t)tr.cur < (uint32_t)RL_TRAV_DONE) trav_node_step<true>(tr, stack, nodes);
        if (tr.cur < 0) blocked = trav_leaf_any(tr, stack, trav) || blocked;
#else
        if ((uint32_t)tr.cur < (uint32_t)RL_TRAV_DONE) trav_node_step<true>(tr, stack, nodes);
        else if (tr.cur < 0) blocked = trav_leaf_any(tr, stack, trav) || blocked;
#endif
        if (tr.cur == RL_TRAV_DONE && my != RL_MISS) {
            if (!blocked && tr.amb) blocked = ref_path_ok(sv, tr.rim_slot, tr.o, tr.d, tr.tmax) || ref_bvh_any(sv, trav, tr.o, tr.d, tr.tmax); // only rim hits block the segment: the reference's boxes decide
            if (!blocked) { // visible: add the light-sampling contribution to the path's accumulator
                float4 c = sh_c[my];
                uint32_t pid = f2u(sh_a[my].w);
                float4 l = lacc[pid];
                l.x += c.x, l.y += c.y, l.z += c.z;
                lacc[pid] = l;
                c_vis++;
            }
            my = RL_MISS;
        }
    }
    for (int off = 16; off > 0; off >>= 1) c_vis += __shfl_down_sync(0xffffffffu, c_vis, off);
    if (lane == 0 && c_vis) atomicAdd(&counters->shadow_visible, (unsigned long long)c_vis);
}

// ---- per-pixel accumulation in sample order (Bitmap::accumulate, structure.rs:397-402) ------------
__global__ void __launch_bounds__(kBlock) k_accum(const float4 *__restrict__ lacc, uint32_t npix, uint32_t n_samples, uint32_t n_slots,
                                                  float4 *__restrict__ img_sum, int first_batch) {
    const size_t n_paths = (size_t)npix * n_samples;
    for (uint32_t lp = blockIdx.x * blockDim.x + threadIdx.x; lp < npix; lp += gridDim.x * blockDim.x) {
        float4 s = first_batch ? make_float4(0.0f, 0.0f, 0.0f, 0.0f) : img_sum[lp];
        for (uint32_t k = 0; k < n_samples; k++) {
            // one pixel sample: its slots in the order of the reference's `l_i +=` statements, then
            // Bitmap::accumulate of the sample (structure.rs:397-402)
            float4 c = lacc[(size_t)k * npix + lp];
            for (uint32_t sl = 1; sl < n_slots; sl++) {
                float4 l = lacc[sl * n_paths + (size_t)k * npix + lp];
                c.x += l.x, c.y += l.y, c.z += l.z;
            }
            s.x += c.x, s.y += c.y, s.z += c.z;
        }
        img_sum[lp] = s;
    }
}
// im_block.scale(1/spp) (integrators/mod.rs:436) + scatter into the full frame
__global__ void __launch_bounds__(kBlock) k_finish(const float4 *__restrict__ img_sum, const uint32_t *__restrict__ pixel_list, uint32_t npix,
                                                   float inv_spp, float *__restrict__ out_rgb) {
    for (uint32_t lp = blockIdx.x * blockDim.x + threadIdx.x; lp < npix; lp += gridDim.x * blockDim.x) {
        float4 s = img_sum[lp];
        uint32_t pixel = pixel_list[lp];
        out_rgb[3 * (size_t)pixel + 0] = s.x * inv_spp;
        out_rgb[3 * (size_t)pixel + 1] = s.y * inv_spp;
        out_rgb[3 * (size_t)pixel + 2] = s.z * inv_spp;
    }
}

__global__ void k_set_u32(uint32_t *p, uint32_t v) { *p = v; }

// End of a batch: what did the wavefront kernels trace?  Queue k was traced by them iff k < *done_at (queue *done_at went to k_tail).
__global__ void k_tally(const uint32_t *__restrict__ qc, const uint32_t *__restrict__ shc, uint32_t n_iters, const uint32_t *__restrict__ done_at, uint32_t iter_offset,
                        Counters *counters) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    unsigned long long seg = 0, sh = 0, deepest = 0;
    const uint32_t lim = min(n_iters, *done_at);
    for (uint32_t k = 0; k < lim; k++) {
        seg += qc[k];
        if (shc) sh += shc[k];
        else if (k == 1) sh += qc[2]; // `direct`: d_counts = {primary rays, extension rays, shadow segments, done_at}
        if (qc[k]) deepest = iter_offset + k + 1;
    }
    counters->wave_segments += seg;
    counters->shadow_traced += sh;
    if (deepest > counters->wave_iters) counters->wave_iters = deepest;
}

// ---- Acceleration::{trace, visible} on caller-provided rays (rl_trace / rl_visible) ---------------
__global__ void __launch_bounds__(kBlock) k_pack_rays(const float *__restrict__ o, const float *__restrict__ d, uint32_t n, float4 *ray_o, float4 *ray_d) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        ray_o[i] = make_float4(o[3 * i], o[3 * i + 1], o[3 * i + 2], u2f(i));
        ray_d[i] = make_float4(d[3 * i], d[3 * i + 1], d[3 * i + 2], 1.0f);
    }
}
__global__ void __launch_bounds__(kBlock) k_primary_rays(SceneView sv, uint32_t w, uint32_t h, float4 *ray_o, float4 *ray_d) {
    uint32_t n = w * h;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint32_t px = i % w, py = i / w;
        V3 o, d;
        camera_generate(sv, (float)px + 0.5f, (float)py + 0.5f, &o, &d);
        ray_o[i] = make_float4(o.x, o.y, o.z, u2f(i));
        ray_d[i] = make_float4(d.x, d.y, d.z, 1.0f);
    }
}
__global__ void __launch_bounds__(kBlock) k_visible_batch(SceneView sv, const float *__restrict__ p0, const float *__restrict__ p1, uint32_t n,
                                                          unsigned char *out) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        V3 a = V3{p0[3 * i], p0[3 * i + 1], p0[3 * i + 2]}, b = V3{p1[3 * i], p1[3 * i + 1], p1[3 * i + 2]};
        out[i] = trace_visible(sv, sv.nodes, sv.trav, a, b) ? 1 : 0;
    }
}

// ---- LBVH build -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_morton(const float4 *__restrict__ verts, uint32_t ntris, V3 smin, V3 sinv, uint64_t *keys) {
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < ntris; p += gridDim.x * blockDim.x) {
        V3 lo, hi;
        tri_bounds(verts, p, &lo, &hi);
        keys[p] = morton_key(lo, hi, smin, sinv, p);
    }
}
__global__ void __launch_bounds__(kBlock) k_tri_setup(const float4 *__restrict__ verts, const uint64_t *__restrict__ keys, uint32_t ntris, float box_eps, float4 *trav,
                                                      float4 *shade, float4 *leaf_lo, float4 *leaf_hi) {
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < ntris; s += gridDim.x * blockDim.x) {
        uint32_t prim = (uint32_t)(keys[s] & 0xffffffffull);
        tri_setup(verts, prim, s, box_eps, trav, shade);
        V3 lo, hi;
        tri_bounds_inflated(verts, prim, box_eps, &lo, &hi);
        leaf_lo[s] = make_float4(lo.x, lo.y, lo.z, 0.0f);
        leaf_hi[s] = make_float4(hi.x, hi.y, hi.z, 0.0f);
    }
}
__global__ void __launch_bounds__(kBlock) k_karras(const uint64_t *__restrict__ keys, int ntris, int2 *children, int2v *ranges, int *parent_of_node,
                                                   int *parent_of_leaf) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ntris - 1; i += gridDim.x * blockDim.x) {
        int l, r, first, last;
        karras_node(keys, ntris, i, &l, &r, &first, &last);
        children[i] = make_int2(l, r);
        ranges[i].x = first;
        ranges[i].y = last;
        if (l < 0) parent_of_leaf[~l] = i;
        else parent_of_node[l] = i;
        if (r < 0) parent_of_leaf[~r] = i;
        else parent_of_node[r] = i;
        if (i == 0) parent_of_node[0] = -1;
    }
}
// Bottom-up fit: the second thread to reach a node owns it (atomic flag), merges the two child
// boxes, writes the wide node (children with at most leaf_max triangles become leaf references)
// and continues to the parent.
__global__ void __launch_bounds__(kBlock) k_fit(int ntris, int leaf_max, const int2 *__restrict__ children, const int2v *__restrict__ ranges,
                                                const int *__restrict__ parent_of_node, const int *__restrict__ parent_of_leaf,
                                                const float4 *__restrict__ leaf_lo, const float4 *__restrict__ leaf_hi, float4 *node_lo, float4 *node_hi,
                                                int *flags, float4 *nodes) {
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < ntris; s += gridDim.x * blockDim.x) {
        int node = parent_of_leaf[s];
        while (node >= 0) {
            if (atomicAdd(&flags[node], 1) == 0) break; // first arrival: the sibling subtree is not ready yet
            __threadfence();
            int2 ch = children[node];
            // boxes written by other SMs: read through L2 (__ldcg), L1 may hold a stale line
            float4 lo0 = ch.x < 0 ? leaf_lo[~ch.x] : __ldcg(&node_lo[ch.x]), hi0 = ch.x < 0 ? leaf_hi[~ch.x] : __ldcg(&node_hi[ch.x]);
            float4 lo1 = ch.y < 0 ? leaf_lo[~ch.y] : __ldcg(&node_lo[ch.y]), hi1 = ch.y < 0 ? leaf_hi[~ch.y] : __ldcg(&node_hi[ch.y]);
            write_wide_node(nodes, node, xyz(lo0), xyz(hi0), xyz(lo1), xyz(hi1), make_child_ref(ch.x, ranges, leaf_max),
                            make_child_ref(ch.y, ranges, leaf_max));
            node_lo[node] = make_float4(fminf(lo0.x, lo1.x), fminf(lo0.y, lo1.y), fminf(lo0.z, lo1.z), 0.0f);
            node_hi[node] = make_float4(fmaxf(hi0.x, hi1.x), fmaxf(hi0.y, hi1.y), fmaxf(hi0.z, hi1.z), 0.0f);
            __threadfence();
            node = parent_of_node[node];
        }
    }
}

} // namespace rl
