"""Turns the gpurun_out/<tag>_* files written by tools/collect_profiles.sh into the tracked summaries under profiles/.
usage: python tools/summarize_r02.py <tag> [out_prefix]"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
pre = sys.argv[2] if len(sys.argv) > 2 else tag
G = os.path.join(ROOT, "gpurun_out")
OUT = os.path.join(ROOT, "profiles")


def read_ncu_csv(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    return rows[hi], rows[hi + 1:]


def short(name):
    n = name.split("(")[0]
    return n.replace("rl::", "").replace("void ", "")


# ---- launch list -----------------------------------------------------------------------------------------------------------------
hdr, data = read_ncu_csv(os.path.join(G, f"{tag}_launches.csv"))
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = collections.defaultdict(float), collections.Counter()
for r in data:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    v = v / 1e6 if r[ui] == "ns" else v / 1e3 if r[ui] == "us" else v
    tot[short(r[ki])] += v
    cnt[short(r[ki])] += 1
T = sum(tot.values())
b = json.load(open(os.path.join(G, f"{tag}_bench_C2.json")))
sm = b["stage_ms"]
with open(os.path.join(OUT, f"{pre}_launches_summary.md"), "w") as f:
    f.write(f"# {pre}: ncu launch list of `python bench.py --steps 1 --warmup 0 --no-cpu-baseline` (config C2)\n\n")
    f.write("Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv ...` (first 600 launches: scene build, the timed step, "
            "the evented step, the e2e steps ...; cold-cache and serialised, so compare SHARES with the live CUDA-event numbers below).\n\n")
    f.write("| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        f.write(f"| `{k}` | {cnt[k]} | {v:.3f} | {v / T * 100:.1f}% |\n")
    live = sum(sm[k] for k in ("ms_trace", "ms_shade", "ms_shadow", "ms_raygen", "ms_accum", "ms_tail"))
    f.write("\n`k_trace_shadow_flat` = closest-hit rays of iteration k + shadow segments of iteration k-1 in one launch; `k_fix_flat` re-traces the rays "
            "whose answer depends on the reference's BVH visit order (ties, rim hits) right after it.\n")
    f.write(f"\nLive CUDA-event stage times of the timed steps of the same build (profiles/{pre}_bench_C2.json, one event per launch, no synchronisation): "
            + ", ".join(f"{k[3:]} {sm[k]:.2f} ms ({sm[k] / live * 100:.1f}%)" for k in ("ms_trace", "ms_shade", "ms_shadow", "ms_raygen", "ms_accum", "ms_tail")) + ".\n")

# ---- full captures ----------------------------------------------------------------------------------------------------------------
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]
with open(os.path.join(OUT, f"{pre}_ncu_full_summary.md"), "w") as f:
    f.write(f"# {pre}: `ncu --set full --clock-control none --import-source on -k regex:<kernel> -c 3 python tools/profile_step.py 8`\n\n")
    f.write("Workload: cbox 1024x1024, 8 spp in one batch (8.4M paths); the three captured launches are wavefront iterations 1-3 "
            "(camera rays; first bounce + shadow segments of the camera hits; second bounce + shadow segments).\n\n")
    for k, kern in (("trace", "k_trace_shadow_flat"), ("shade", "k_shade")):
        rep = os.path.join(G, f"{tag}_{k}.ncu-rep")
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rr = list(csv.reader(txt.splitlines()))
        h, u = rr[0], rr[1]
        f.write(f"## {kern}\n\n| metric | unit | launch 1 | launch 2 | launch 3 |\n|---|---|---:|---:|---:|\n")
        for w in want:
            if w in h:
                i = h.index(w)
                f.write(f"| {w} | {u[i]} | " + " | ".join(r[i] for r in rr[2:5]) + " |\n")
        f.write("\n")

# ---- executed instructions per source line ----------------------------------------------------------------------------------------
with open(os.path.join(OUT, f"{pre}_by_line.md"), "w") as f:
    f.write(f"# {pre}: executed warp instructions per source line (tools/ncu_by_line.py on the --set full --import-source captures)\n")
    for k, kern, launches in (("trace", "k_trace_shadow_flat", (0, 1)), ("shade", "k_shade", (0,))):
        for li in launches:
            out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_by_line.py"), os.path.join(G, f"{tag}_{k}.ncu-rep"), kern, str(li)],
                                 capture_output=True, text=True).stdout
            f.write(f"\n## {kern}, launch {li + 1}\n```\n{out}```\n")

# ---- DRAM traffic of the dominant kernel at the bench configuration -------------------------------------------------------------------
traffic = {}
for k, kern in (("trace", "k_trace_shadow_flat"), ("shade", "k_shade")):
    hdr, data = read_ncu_csv(os.path.join(G, f"{tag}_traffic_{k}.csv"))
    ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    by_id = collections.defaultdict(dict)
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        unit = r[ui].lower()
        if r[mi].startswith("dram__bytes"):
            v *= {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[unit]
        by_id[r[0]][r[mi]] = v
    launches = list(by_id.values())
    dram = sum(x.get("dram__bytes_read.sum", 0) + x.get("dram__bytes_write.sum", 0) for x in launches)
    traffic[kern] = {"launches": len(launches), "dram_bytes_total": dram, "dram_bytes_per_launch": dram / max(1, len(launches))}
ru = b["roofline"]["units_per_step"]
alg_trace = ru["segments"] * 48 + ru["shadow_segments"] * 80
tj = {"C2": {"dram_bytes_per_algorithmic_byte": traffic["k_trace_shadow_flat"]["dram_bytes_total"] / alg_trace,
             "source": f"profiles/{pre}_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum over ALL {traffic['k_trace_shadow_flat']['launches']} launches of k_trace_shadow_flat "
                       "of one C2 frame (ncu --metrics, tools/profile_step.py 128) / algorithmic bytes of the frame (48 B per segment + 80 B per traced shadow segment)",
             "k_trace_shadow_flat": traffic["k_trace_shadow_flat"], "k_shade": traffic["k_shade"],
             "k_shade_dram_bytes_per_algorithmic_byte": traffic["k_shade"]["dram_bytes_total"] / (b["roofline_shade"]["units_per_step"] * 192)}}
json.dump(tj, open(os.path.join(OUT, "traffic.json"), "w"), indent=1)
json.dump(tj, open(os.path.join(OUT, f"{pre}_traffic.json"), "w"), indent=1)
for name in ("bench_C2.json", "bench_C2_ref.json"):
    open(os.path.join(OUT, f"{pre}_{name}"), "w").write(open(os.path.join(G, f"{tag}_{name}")).read())
print(json.dumps(tj["C2"], indent=1)[:900])
