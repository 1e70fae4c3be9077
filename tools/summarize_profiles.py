"""Turns gpurun_out/{launches_*.csv, prof_*.ncu-rep, bench_*.json} into the tracked summaries under profiles/.
usage: python tools/summarize_profiles.py <tag> <launch_csv> <trace.ncu-rep> <shade.ncu-rep> <shadow.ncu-rep> <bench.json> <bench_ref.json> [note]"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, launches, p_trace, p_shade, p_shadow, bench, bench_ref = sys.argv[1:8]
note = sys.argv[8] if len(sys.argv) > 8 else ""
out = os.path.join(ROOT, "profiles")
os.makedirs(out, exist_ok=True)

rows = list(csv.reader(open(launches)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr, data = rows[hi], rows[hi + 1:]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = collections.defaultdict(float), collections.Counter()
for r in data:
    if len(r) <= vi:
        continue
    name = r[ki].split("(")[0][:60]
    v = float(r[vi].replace(",", ""))
    v = v / 1e6 if r[ui] == "ns" else v / 1e3 if r[ui] == "us" else v
    tot[name] += v
    cnt[name] += 1
T = sum(tot.values())
b = json.load(open(bench))
sm = b["stage_ms"]
with open(os.path.join(out, f"{tag}_launches_summary.md"), "w") as f:
    f.write(f"# {tag}: ncu launch list ({note})\n\n")
    f.write("Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline`\n")
    f.write("(first 1500 launches; cold-cache, serialised, so compare SHARES with the live CUDA-event numbers below).\n\n")
    f.write("| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        f.write(f"| `{k}` | {cnt[k]} | {v:.3f} | {v / T * 100:.1f}% |\n")
    live = sm["ms_trace"] + sm["ms_shade"] + sm["ms_shadow"] + sm["ms_raygen"] + sm["ms_accum"]
    f.write("\n`k_trace_shadow_flat` = closest-hit rays of iteration k+1 + shadow segments of iteration k in one launch: its share is "
            "the sum of the trace and shadow stages below (the profiled step times them as separate kernels).\n")
    f.write(f"\nLive CUDA-event stage times of one profiled step of the same build ({bench}): "
            + ", ".join(f"{k[3:]} {sm[k]:.2f} ms ({sm[k] / live * 100:.1f}%)" for k in ("ms_trace", "ms_shade", "ms_shadow", "ms_raygen", "ms_accum")) + ".\n")

want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum"]
traffic = {}
with open(os.path.join(out, f"{tag}_ncu_full_summary.md"), "w") as f:
    f.write(f"# {tag}: `ncu --set full --clock-control none --import-source on -k regex:<kernel> -c 3 python tools/profile_step.py 8` ({note})\n\n")
    f.write("Workload: cbox 1024x1024, 8 spp in one batch (8.4M paths); the three captured launches are wavefront iterations 1-3 "
            "(camera rays, first and second bounce).\n\n")
    for k, rep in (("trace", p_trace), ("shade", p_shade), ("shadow", p_shadow)):
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rr = list(csv.reader(txt.splitlines()))
        h, u = rr[0], rr[1]
        f.write(f"## k_{k}{'' if k == 'shade' else '_flat'}\n\n| metric | unit | launch 1 | launch 2 | launch 3 |\n|---|---|---:|---:|---:|\n")
        for w in want:
            if w in h:
                i = h.index(w)
                f.write(f"| {w} | {u[i]} | " + " | ".join(r[i] for r in rr[2:5]) + " |\n")
        f.write("\n")
        if k == "trace":
            ir, iw = h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
            scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
            per = [float(r[ir]) * scale[u[ir]] + float(r[iw]) * scale[u[iw]] for r in rr[2:5]]
            traffic = {"k_trace_dram_bytes_per_launch": sum(per) / len(per), "per_launch": per,
                       # launch 1 traces exactly 1024*1024*8 camera rays: measured DRAM bytes per algorithmic byte (48 B / ray)
                       "launch1_rays": 1024 * 1024 * 8, "dram_bytes_per_algorithmic_byte": per[0] / (1024 * 1024 * 8 * 48.0),
                       "source": f"profiles/{tag}_ncu_full_summary.md (dram__bytes_read.sum + dram__bytes_write.sum, mean of the 3 captured k_trace launches; "
                                 "algorithmic bytes of those launches: 8.39M, ~7.0M, ~4.5M rays x 48 B)"}
json.dump(traffic, open(os.path.join(out, "traffic.json"), "w"), indent=1)
for src, dst in ((bench, f"{tag}_bench.json"), (bench_ref, f"{tag}_bench_reference.json")):
    open(os.path.join(out, dst), "w").write(open(src).read())
print(open(os.path.join(out, f"{tag}_launches_summary.md")).read())
