#!/usr/bin/env python
"""bench.py -- headline benchmark: Cornell box `path`, 1024x1024, 128 spp (BASELINE.json configs[1]).

    python bench.py --gpus 1 --steps 5 --warmup 3                      # our arm, one GPU
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   # image-tile split + one NCCL reduce
    python bench.py --impl reference --gpus 1 --steps 1 --warmup 0     # CPU arm: the oracle, all host threads

A step is one rl_render of the whole frame (every sample of every pixel once).  `value` is
Msamples/s with the scene resident in HBM and the result left on the device, timed with CUDA
events inside the library (max over ranks); `e2e` goes through the public call with host
buffers: scene upload + LBVH build + render + framebuffer read-back, every step.
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

WORKLOAD = dict(scene="data/cbox.pbrt", scale_image=2.0, width=1024, height=1024, spp=128, integrator="path",
                min_depth=0, max_depth=None, rr_depth=0, strategy="all", seed=0)
METRIC = "Msamples/sec (Cornell box path, 1024x1024, 128spp); Mpath-segments/sec alongside"
UNIT = "Msamples/s"
# closest-hit traversal reads one 32-byte ray and writes one 16-byte hit per segment
TRACE_BYTES_PER_SEGMENT = 48
SHADE_BYTES_PER_VERTEX = 192  # 64 R + 48 W + 48 W (shadow segment) + 32 RMW


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


def load_scene():
    from rustlight_b200 import SceneLoaderManager
    sc = SceneLoaderManager().load(os.path.join(ROOT, WORKLOAD["scene"]))
    sc.scale_image(WORKLOAD["scale_image"])  # CLI `-s 2`
    assert sc.size == (WORKLOAD["width"], WORKLOAD["height"])
    return sc


def integ_desc():
    from rustlight_b200 import _abi
    return _abi.path_desc(WORKLOAD["min_depth"], WORKLOAD["max_depth"], WORKLOAD["rr_depth"], _abi.RL_STRATEGY_ALL, False)


def cpu_reference_run(spp, nthreads=0):
    """The reference's own CPU path as restated by the oracle: graph estimator, BVHAccel, glibc
    math, per-block xoshiro streams (mode A), std::thread over 16x16 blocks.  Timed region =
    block loop + merge, like the reference's "Elapsed Integrator" (integrators/mod.rs:323-334)."""
    from oracle import binding as ob
    from rustlight_b200 import _abi
    sc = load_scene()
    osc = ob.OracleScene(sc)
    cfg = ob.config(math_mode=ob.MATH_LIBM, accel_mode=ob.ACCEL_BVH, estimator=ob.EST_GRAPH, nthreads=nthreads)
    _, st = osc.render(integ_desc(), spp, seed=WORKLOAD["seed"], sampler_mode=_abi.RL_SAMPLER_BLOCK_STREAM, cfg=cfg)
    return st


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    spp = args.ref_spp
    for _ in range(args.warmup):
        cpu_reference_run(1)
    secs, samples, segs, cores = 0.0, 0, 0, 0
    for _ in range(args.steps):
        st = cpu_reference_run(spp)
        secs += st.seconds
        samples += st.samples
        segs += st.segments
        cores = st.threads_used
    v = samples / secs / 1e6
    sample = f"{WORKLOAD['width']}x{WORKLOAD['height']} x {spp} spp per step (of 128), sampler mode A, C++ restatement of rustlight@864df34"
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": "cbox path 1024x1024 128spp (bounded CPU sample)", **{k: WORKLOAD[k] for k in ("integrator", "strategy", "rr_depth")}},
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
    ap.add_argument("--ref-spp", type=int, default=32, help="spp of the bounded CPU sample per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--batch-spp", type=int, default=0)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

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

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    import numpy as np
    ctx = Context(local_rank, nranks=world, rank=rank, nccl_id=nccl_id)
    scene = load_scene()
    dsc = DeviceScene(ctx, scene)
    integ = integ_desc()
    spp, seed = WORKLOAD["spp"], WORKLOAD["seed"]
    W, H = scene.size
    from rustlight_b200.device import PinnedImage
    pinned = PinnedImage(H, W)  # the step's result is read back into pinned host memory (rl_host_alloc)
    out = pinned.array
    from rustlight_b200 import _abi
    import ctypes as C
    from rustlight_b200.device import lib
    L = lib()
    opts = _abi.rl_render_opts(C.sizeof(_abi.rl_render_opts), spp, seed, _abi.RL_SAMPLER_COUNTER, args.batch_spp, 2, 0)  # material_sort = auto (off: one BSDF kind)
    FP = C.POINTER(C.c_float)

    def step_device():
        st = _abi.rl_stats()
        # result stays on the device; with world > 1 the single ncclReduce is part of the step
        ctx._check(L.rl_render(ctx._h, dsc._h, C.byref(integ), C.byref(opts), None, C.byref(st)))
        return st

    for _ in range(max(args.warmup, 0)):
        step_device()
    clocks = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        clocks.start()
    t0 = time.perf_counter()
    dev_ms, segs, launches, samples, shadows = 0.0, 0, 0, 0, 0
    for _ in range(args.steps):
        st = step_device()
        dev_ms += st.ms_total + st.ms_reduce
        segs += st.segments
        shadows += st.shadow_rays
        launches += st.kernel_launches
        samples += st.samples
    wall_ms = (time.perf_counter() - t0) * 1e3
    barrier()
    clk = clocks.stop() if rank == 0 else None
    dev_ms = max_over_ranks(dev_ms)
    wall_ms = max_over_ranks(wall_ms)
    tot_samples = sum_over_ranks(samples)
    tot_segs = sum_over_ranks(segs)
    tot_shadows = sum_over_ranks(shadows)
    tot_launches = sum_over_ranks(launches)
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

    # ---- roofline of the dominant kernel (closest-hit traversal), per-stage CUDA events ----
    roof = None
    roof_shade = None
    stage_ms = None
    # every rank runs the profiled step: rl_render contains the collective (ncclReduce)
    barrier()
    ctx.set_profiling(True)
    st = step_device()
    ctx.set_profiling(False)
    barrier()
    if rank == 0:
        peak, peak_src = measured_peaks()
        # per profiled step: 1 k_finish + per batch (1 k_raygen + 1 k_set_u32-free count + 1 k_accum) + 3 kernels per wavefront iteration
        launches_trace = max(1, (int(st.kernel_launches) - 1) // 3)
        bytes_total = st.segments * TRACE_BYTES_PER_SEGMENT
        achieved = bytes_total / (st.ms_trace * 1e-3) / 1e9 if st.ms_trace > 0 else 0.0
        stage_ms = {k: getattr(st, k) for k in ("ms_total", "ms_raygen", "ms_trace", "ms_shade", "ms_shadow", "ms_accum")}
        traffic, traffic_src = None, None
        tp = os.path.join(ROOT, "profiles", "traffic.json")  # dram bytes per k_trace launch from the committed ncu --set full capture
        if os.path.exists(tp):
            tj = json.load(open(tp))
            # the capture is of a smaller render (8 spp); its launch 1 traces exactly 8 Mi camera rays, which gives the measured
            # DRAM bytes per algorithmic byte of the kernel; scaled to this run's average launch
            ratio = tj.get("dram_bytes_per_algorithmic_byte")
            if ratio is not None:
                traffic = ratio * bytes_total / launches_trace
                traffic_src = tj.get("source", "") + f"; {ratio:.3f} DRAM bytes per algorithmic byte, scaled to this run's launch size"
            else:
                traffic, traffic_src = tj.get("k_trace_dram_bytes_per_launch"), tj.get("source")
        roof = {"kernel": "k_trace_flat (closest hit: group-table scan + exact triangle tests)", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_unit": TRACE_BYTES_PER_SEGMENT, "units": "path segments", "units_per_step": int(st.segments),
                "launches_per_step": launches_trace, "avg_launch_ms": st.ms_trace / launches_trace,
                "share_of_step": st.ms_trace / st.ms_total if st.ms_total else None,
                "algorithmic_bytes_per_launch": bytes_total / launches_trace,
                "note": "group table + triangle records (5 KB) are shared-memory resident: the kernel is instruction-issue bound, not HBM bound (SURVEY.md F9; profiles/). "
                        "Timed alone (k_trace_flat) in the profiled step; the timed steps run it inside k_trace_shadow_flat together with the shadow segments "
                        "of the previous iteration (one launch instead of two), so the ncu launch list shows that kernel with the sum of both shares"}

        # the second large stage, same method (north_star: "traversal and shade kernels"): one surface vertex = 64 B read (ray, state,
        # hit) + 48 B next ray/state + 48 B shadow segment + 32 B accumulator RMW = 192 B
        shade_bytes = st.hits * SHADE_BYTES_PER_VERTEX
        shade_achieved = shade_bytes / (st.ms_shade * 1e-3) / 1e9 if st.ms_shade > 0 else 0.0
        roof_shade = {"kernel": "k_shade (surface interaction, BSDF sample + RR, light sample)", "bound": "hbm", "achieved": shade_achieved, "peak": peak,
                      "unit": "GB/s", "frac": shade_achieved / peak, "algorithmic_bytes_per_unit": SHADE_BYTES_PER_VERTEX,
                      "units": "surface vertices (hits)", "units_per_step": int(st.hits), "launches_per_step": launches_trace,
                      "avg_launch_ms": st.ms_shade / launches_trace, "share_of_step": st.ms_shade / st.ms_total if st.ms_total else None}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        stc = cpu_reference_run(args.ref_spp)
        cpu = {"value": stc.samples / stc.seconds / 1e6, "unit": UNIT, "cores": int(stc.threads_used), "kind": "port",
               "mpath_segments_per_s": stc.segments / stc.seconds / 1e6,
               "sample": f"1024x1024 x {args.ref_spp} spp (of 128), {stc.seconds:.1f} s; C++ restatement of rustlight@864df34 "
                         "(graph estimator, BVHAccel, glibc math, per-block xoshiro streams)"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": "cbox.pbrt path -n 128 -s 2 (1024x1024), independent:0 -> counter stream (mode B)",
                           "integrator": "path", "strategy": "all", "rr_depth": 0, "max_depth": "inf", "partition": f"16x16 tiles over {world} rank(s), 1 ncclReduce",
                           "l2": "wavefront queues are 23.6 GB per batch (176 B x 134 M paths in flight), far larger than the 126 MB L2"},
                "mpath_segments_per_s": tot_segs / (dev_ms * 1e-3) / 1e6,
                "mshadow_rays_per_s": tot_shadows / (dev_ms * 1e-3) / 1e6,  # Acceleration::visible calls of the reference
                "wall_ms_per_step": wall_ms / args.steps,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms / args.steps},
                "gpu_launches": int(tot_launches), "clocks": clk, "roofline": roof, "roofline_shade": roof_shade, "stage_ms": stage_ms, "cpu_baseline": cpu}
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
