#!/usr/bin/env python
"""bench.py -- the five BASELINE.json configurations; the headline (default) is C2: Cornell box `path`, 1024x1024, 128 spp.

    python bench.py --gpus 1 --steps 5 --warmup 3                      # our arm, one GPU, config C2
    python bench.py --config C4 ...                                    # C1..C5 (BASELINE.json configs[0..4])
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   # image-tile split + one NCCL reduce
    python bench.py --impl reference --gpus 1 --steps 1 --warmup 0     # CPU arm: the oracle, all host threads

A step is one rl_render of the whole frame (every sample of every pixel once).  `value` is Msamples/s with the scene
resident in HBM and the result left on the device, timed with CUDA events inside the library (max over ranks); `e2e` goes
through the public call with host buffers: scene upload + LBVH build + render + framebuffer read-back, every step.
`roofline` is measured in the timed steps themselves: the library records one CUDA event after every launch (no
synchronisation) and the durations of the dominant kernel's launches are summed when the frame is complete.
The oracle is imported only for the cpu_baseline / --impl reference legs.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WALL_KD = [(0.63, 0.065, 0.05), (0.14, 0.45, 0.091), (0.725, 0.71, 0.68)]
# BASELINE.json configs[0..4].  ref_spp: samples per pixel of the bounded CPU sample (cpu_baseline / --impl reference):
# the same frame at fewer samples; throughput per sample does not depend on the sample count (every sample is independent).
CONFIGS = {
    "C1": dict(label="cbox.pbrt path -n 16 (512x512)", width=512, height=512, spp=16, integrator="path", ref_spp=16, gpus=1),
    "C2": dict(label="cbox.pbrt path -n 128 -s 2 (1024x1024)", width=1024, height=1024, spp=128, integrator="path", ref_spp=32, gpus=1),
    "C3": dict(label="cbox.pbrt with Phong walls (kd = 0.5 x wall colour, ks = 0.3, exponent 50) path -n 512 (512x512)", width=512, height=512,
               spp=512, integrator="path", phong_walls=True, ref_spp=32, gpus=1),
    "C4": dict(label="cbox.pbrt direct -b 1 -l 1 -n 64 -s 4 (2048x2048)", width=2048, height=2048, spp=64, integrator="direct", ref_spp=8, gpus=8),
    "C5": dict(label="cbox.pbrt path -n 4096, Film 1920x1080, per-bounce material sort on", width=1920, height=1080, spp=4096, integrator="path",
               material_sort=1, ref_spp=8, gpus=8),
}
UNIT = "Msamples/s"
# closest-hit traversal reads one 32-byte ray and writes one 16-byte hit per segment; a shadow segment costs
# 36 B of traversal (32 R + 4 W) + 44 B of light-sample resolve (4 R + 16 R + 24 RMW)
TRACE_BYTES_PER_SEGMENT = 48
SHADOW_BYTES_PER_SEGMENT = 80
SHADE_BYTES_PER_VERTEX = 192  # 64 R + 48 W + 48 W (shadow segment) + 32 RMW


def metric_name(cfg):
    return f"Msamples/sec (Cornell box {cfg['integrator']}, {cfg['width']}x{cfg['height']}, {cfg['spp']}spp); Mpath-segments/sec alongside"


def workload_string(cfg):
    return cfg["label"] + ", independent:0 -> counter stream (mode B) on the GPU / per-block xoshiro streams (mode A) on the CPU"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index=0):
        self.gpu_index = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])), mx.append(float(f[2])), pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        load = [s for s, p in zip(sm, pw) if p >= 0.5 * max(pw)] or sm
        return {"sm_mhz": statistics.median(load), "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


def load_scene(cfg):
    from rustlight_b200 import SceneLoaderManager
    from rustlight_b200.host import material_phong
    sc = SceneLoaderManager().load(os.path.join(ROOT, "data", "cbox.pbrt"))
    if (cfg["width"], cfg["height"]) != sc.size:
        if cfg["width"] == cfg["height"] and cfg["width"] % sc.size[0] == 0:
            sc.scale_image(cfg["width"] / sc.size[0])  # CLI `-s`
        else:
            sc.set_resolution(cfg["width"], cfg["height"])  # edited Film (note the Fov::Y quirk, camera.rs:41-44)
    assert sc.size == (cfg["width"], cfg["height"])
    if cfg.get("phong_walls"):
        for mesh, kd in [(0, WALL_KD[2]), (1, WALL_KD[2]), (2, WALL_KD[2]), (3, WALL_KD[1]), (4, WALL_KD[0])]:
            sc.set_material(mesh, material_phong([0.5 * c for c in kd], (0.3, 0.3, 0.3), 50.0))
    return sc


def integ_desc(cfg):
    from rustlight_b200 import _abi
    if cfg["integrator"] == "direct":
        return _abi.direct_desc(1, 1)
    return _abi.path_desc(0, None, 0, _abi.RL_STRATEGY_ALL, False)  # CLI defaults: min 0, max inf, rr 0, strategy all


def cpu_reference_run(cfg, spp, nthreads=0):
    """The reference's own CPU path as restated by the oracle: graph estimator, BVHAccel, glibc
    math, per-block xoshiro streams (mode A), std::thread over 16x16 blocks.  Timed region =
    block loop + merge, like the reference's "Elapsed Integrator" (integrators/mod.rs:323-334)."""
    from oracle import binding as ob
    from rustlight_b200 import _abi
    osc = ob.OracleScene(load_scene(cfg))
    ocfg = ob.config(math_mode=ob.MATH_LIBM, accel_mode=ob.ACCEL_BVH, estimator=ob.EST_GRAPH, nthreads=nthreads)
    _, st = osc.render(integ_desc(cfg), spp, seed=0, sampler_mode=_abi.RL_SAMPLER_BLOCK_STREAM, cfg=ocfg)
    return st


def config_block(cfg, name, world):
    return {"workload": workload_string(cfg), "config": name, "integrator": cfg["integrator"], "strategy": "all", "rr_depth": 0, "max_depth": "inf",
            "partition": f"16x16 tiles over {world} rank(s), 1 ncclReduce",
            "cpu_sample": f"the CPU arm (cpu_baseline, --impl reference) renders the same frame at {cfg['ref_spp']} of {cfg['spp']} spp per step; "
                          "throughput is per sample, samples are independent",
            "l2": "inputs larger than L2: the wavefront queues of a batch are 176 B per path in flight (C2: 23.6 GB), far beyond the 126 MB L2"}


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    spp = args.ref_spp or cfg["ref_spp"]
    for _ in range(args.warmup):
        cpu_reference_run(cfg, 1)
    secs, samples, segs, cores = 0.0, 0, 0, 0
    for _ in range(args.steps):
        st = cpu_reference_run(cfg, spp)
        secs += st.seconds
        samples += st.samples
        segs += st.segments
        cores = st.threads_used
    v = samples / secs / 1e6
    sample = f"{cfg['width']}x{cfg['height']} x {spp} spp per step (of {cfg['spp']}), sampler mode A, C++ restatement of rustlight@864df34"
    line = {"impl": "reference", "metric": metric_name(cfg), "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": config_block(cfg, args.config, 1),
            "mpath_segments_per_s": segs / secs / 1e6,
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)
    return 0


_RESULT = None  # the process's real stdout, kept for the one JSON line


def claim_stdout():
    """Exactly ONE line may reach stdout (the contract): point fd 1 at stderr for everything libraries print there
    (NCCL writes its version banner to stdout on communicator creation) and keep the real stdout for the result line."""
    global _RESULT
    if _RESULT is None:
        sys.stdout.flush()
        _RESULT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
    return _RESULT


def emit(line):
    out = claim_stdout()
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C2", choices=sorted(CONFIGS))
    ap.add_argument("--ref-spp", type=int, default=0, help="spp of the bounded CPU sample per step (0 = the config's default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--batch-spp", type=int, default=0)
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        return run_reference(args, cfg)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    from rustlight_b200.device import Context, DeviceScene, nccl_unique_id
    dist = None
    nccl_id = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        ids = [nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        nccl_id = ids[0]

    def barrier():
        if dist is not None:
            dist.barrier()

    def reduce_ranks(x, op):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return float(t.item())

    def max_over_ranks(x):
        return reduce_ranks(x, dist.ReduceOp.MAX) if dist is not None else x

    def min_over_ranks(x):
        return reduce_ranks(x, dist.ReduceOp.MIN) if dist is not None else x

    def sum_over_ranks(x):
        return reduce_ranks(x, dist.ReduceOp.SUM) if dist is not None else x

    ctx = Context(local_rank, nranks=world, rank=rank, nccl_id=nccl_id)
    scene = load_scene(cfg)
    dsc = DeviceScene(ctx, scene)
    integ = integ_desc(cfg)
    spp, seed = cfg["spp"], 0
    W, H = scene.size
    from rustlight_b200.device import PinnedImage
    pinned = PinnedImage(H, W)  # the step's result is read back into pinned host memory (rl_host_alloc)
    out = pinned.array
    from rustlight_b200 import _abi
    import ctypes as C
    from rustlight_b200.device import lib
    L = lib()
    opts = _abi.rl_render_opts(C.sizeof(_abi.rl_render_opts), spp, seed, _abi.RL_SAMPLER_COUNTER, args.batch_spp, cfg.get("material_sort", 2), 0)
    FP = C.POINTER(C.c_float)

    def step_device():
        st = _abi.rl_stats()
        # result stays on the device; with world > 1 the single ncclReduce is part of the step
        ctx._check(L.rl_render(ctx._h, dsc._h, C.byref(integ), C.byref(opts), None, C.byref(st)))
        return st

    for _ in range(max(args.warmup, 0)):
        step_device()

    def timed_steps():
        a = dict(ms=0.0, segs=0, launches=0, samples=0, nee=0, sh_traced=0, hits=0, ms_trace=0.0, ms_shade=0.0, ms_tail=0.0, ms_shadow=0.0,
                 ms_raygen=0.0, ms_accum=0.0, l_trace=0, l_shade=0)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            st = step_device()
            a["ms"] += st.ms_total + st.ms_reduce
            a["segs"] += st.segments
            a["nee"] += st.shadow_rays
            a["sh_traced"] += st.shadow_traced
            a["hits"] += st.hits
            a["launches"] += st.kernel_launches
            a["samples"] += st.samples
            for k in ("ms_trace", "ms_shade", "ms_tail", "ms_shadow", "ms_raygen", "ms_accum"):
                a[k] += getattr(st, k)
            a["l_trace"] += st.launches_trace
            a["l_shade"] += st.launches_shade
        a["wall_ms"] = (time.perf_counter() - t0) * 1e3
        barrier()
        return a

    # ---- the timed region: K steps, device-timed inside the library, max over ranks ----
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    plain = timed_steps()
    # ---- the same K steps again with one CUDA event after every launch (no synchronisation, the same kernels): per-kernel durations for the
    # roofline.  The events cost host time per launch, which shows at small frames / many ranks, so `value` comes from the steps above.
    ctx.set_profiling(2)
    acc = timed_steps()
    ctx.set_profiling(0)
    clk = clocks.stop() if rank == 0 else None
    my_ms = plain["ms"] / args.steps
    dev_ms = max_over_ranks(plain["ms"])
    evented_ms = max_over_ranks(acc["ms"])
    rank_ms = {"min": min_over_ranks(my_ms), "max": max_over_ranks(my_ms), "mean": sum_over_ranks(my_ms) / world}
    wall_ms = max_over_ranks(plain["wall_ms"])
    tot_samples = sum_over_ranks(plain["samples"])
    tot_segs = sum_over_ranks(plain["segs"])
    tot_nee = sum_over_ranks(plain["nee"])
    tot_sh = sum_over_ranks(plain["sh_traced"])
    tot_launches = sum_over_ranks(plain["launches"])
    value = tot_samples / (dev_ms * 1e-3) / 1e6

    # ---- e2e: public call with host buffers, scene upload + build + render + read-back every step ----
    def step_e2e():
        d = DeviceScene(ctx, scene)
        st = _abi.rl_stats()
        ptr = out.ctypes.data_as(FP)
        ctx._check(L.rl_render(ctx._h, d._h, C.byref(integ), C.byref(opts), ptr, C.byref(st)))
        d.close()
        return st
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    barrier()
    e2e_ms = max_over_ranks(e2e_ms)
    e2e_value = tot_samples / (e2e_ms * 1e-3) / 1e6
    desc = scene.desc.contents
    h2d = sum(desc.meshes[i].nverts * (12 + 12 + 8) + desc.meshes[i].ntris * 12 for i in range(desc.nmeshes)) + 136 + 32 + 32
    d2h = W * H * 12 if rank == 0 else 0

    # ---- roofline of the dominant kernel, from the events of the timed steps (rank 0's launches) ----
    roof = roof_shade = roof_stages = stage_ms = None
    # per-stage kernels (trace and shadow launched separately) for the supplementary per-stage numbers: every rank runs the step (rl_render holds the collective)
    barrier()
    ctx.set_profiling(1)
    st1 = step_device()
    ctx.set_profiling(0)
    barrier()
    if rank == 0:
        peak, peak_src = measured_peaks()
        path = cfg["integrator"] == "path"
        tj = {}
        tp = os.path.join(ROOT, "profiles", "traffic.json")  # DRAM bytes per algorithmic byte of the kernel from the committed ncu capture
        if os.path.exists(tp):
            tj = json.load(open(tp)).get(args.config, {})
        if path:
            name = "k_trace_shadow_flat (closest hit of iteration k + shadow segments of iteration k-1 in one launch: quad-table scan + exact triangle tests)"
            alg = acc["segs"] * TRACE_BYTES_PER_SEGMENT + acc["sh_traced"] * SHADOW_BYTES_PER_SEGMENT
        else:
            name = "k_trace_flat (closest hit: quad-table scan + exact triangle tests)"
            alg = acc["segs"] * TRACE_BYTES_PER_SEGMENT
        launches = max(1, acc["l_trace"])
        achieved = alg / (acc["ms_trace"] * 1e-3) / 1e9 if acc["ms_trace"] > 0 else 0.0
        ratio = tj.get("dram_bytes_per_algorithmic_byte")
        roof = {"kernel": name, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ratio * alg / launches if ratio is not None else None, "traffic_source": tj.get("source"), "peak_source": peak_src,
                "algorithmic_bytes": f"{TRACE_BYTES_PER_SEGMENT} B per closest-hit segment" + (f" + {SHADOW_BYTES_PER_SEGMENT} B per traced shadow segment" if path else ""),
                "units_per_step": {"segments": acc["segs"] // args.steps, "shadow_segments": acc["sh_traced"] // args.steps},
                "launches_per_step": launches / args.steps, "avg_launch_ms": acc["ms_trace"] / launches,
                "share_of_step": acc["ms_trace"] / acc["ms"] if acc["ms"] else None, "algorithmic_bytes_per_launch": alg / launches,
                "measured": "CUDA events around every launch of the timed steps (rl_set_profiling 2: no synchronisation, the kernels of an untimed frame)",
                "note": "quad table + triangle records (5 KB) are shared-memory resident: the kernel is instruction-issue bound, not HBM bound (SURVEY.md F9; profiles/)"}
        shade_bytes = acc["hits"] * SHADE_BYTES_PER_VERTEX
        l_sh = max(1, acc["l_shade"])
        sh_ach = shade_bytes / (acc["ms_shade"] * 1e-3) / 1e9 if acc["ms_shade"] > 0 else 0.0
        roof_shade = {"kernel": "k_shade (surface interaction, BSDF sample + RR, light sample)" if path else "k_shade_direct1/2", "bound": "hbm", "achieved": sh_ach,
                      "peak": peak, "unit": "GB/s", "frac": sh_ach / peak, "algorithmic_bytes_per_unit": SHADE_BYTES_PER_VERTEX,
                      "units": "surface vertices (hits)", "units_per_step": acc["hits"] // args.steps, "launches_per_step": l_sh / args.steps,
                      "avg_launch_ms": acc["ms_shade"] / l_sh, "share_of_step": acc["ms_shade"] / acc["ms"] if acc["ms"] else None}
        stage_ms = {k: acc[k] / args.steps for k in ("ms_raygen", "ms_trace", "ms_shade", "ms_shadow", "ms_tail", "ms_accum")}
        stage_ms["ms_total"] = acc["ms"] / args.steps

        def frac(b, ms):
            return (b / (ms * 1e-3) / 1e9) / peak if ms > 0 else None
        roof_stages = {"source": "one extra step with trace and shadow launched as separate kernels (rl_set_profiling 1), same events",
                       "k_trace_flat": {"ms": st1.ms_trace, "frac": frac(st1.segments * TRACE_BYTES_PER_SEGMENT, st1.ms_trace), "bytes_per_segment": TRACE_BYTES_PER_SEGMENT},
                       "k_shadow_flat": {"ms": st1.ms_shadow, "frac": frac(st1.shadow_traced * SHADOW_BYTES_PER_SEGMENT, st1.ms_shadow), "bytes_per_segment": SHADOW_BYTES_PER_SEGMENT},
                       "k_shade": {"ms": st1.ms_shade, "frac": frac(st1.hits * SHADE_BYTES_PER_VERTEX, st1.ms_shade), "bytes_per_vertex": SHADE_BYTES_PER_VERTEX},
                       "ms_tail": st1.ms_tail, "ms_total": st1.ms_total}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rspp = args.ref_spp or cfg["ref_spp"]
        stc = cpu_reference_run(cfg, rspp)
        cpu = {"value": stc.samples / stc.seconds / 1e6, "unit": UNIT, "cores": int(stc.threads_used), "kind": "port",
               "mpath_segments_per_s": stc.segments / stc.seconds / 1e6,
               "sample": f"{W}x{H} x {rspp} spp (of {spp}), {stc.seconds:.1f} s; C++ restatement of rustlight@864df34 "
                         "(graph estimator, BVHAccel, glibc math, per-block xoshiro streams)"}

    if rank == 0:
        line = {"metric": metric_name(cfg), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": config_block(cfg, args.config, world),
                "mpath_segments_per_s": tot_segs / (dev_ms * 1e-3) / 1e6,
                "mvisible_calls_per_s": tot_nee / (dev_ms * 1e-3) / 1e6,  # Acceleration::visible calls of the reference, parity counter
                "mshadow_rays_per_s": tot_sh / (dev_ms * 1e-3) / 1e6,     # shadow segments actually traced
                "wall_ms_per_step": wall_ms / args.steps, "ms_per_step_with_launch_events": evented_ms / args.steps,
                "rank_ms_per_step": rank_ms,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms / args.steps},
                "gpu_launches": int(tot_launches), "clocks": clk, "roofline": roof, "roofline_shade": roof_shade, "roofline_stages": roof_stages,
                "stage_ms": stage_ms, "cpu_baseline": cpu}
        emit(line)
    dsc.close()
    out = None
    pinned.close()
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
