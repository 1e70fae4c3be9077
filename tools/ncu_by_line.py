"""Join an ncu SASS-level source page with nvdisasm line info: executed warp instructions per source line.

usage: python tools/ncu_by_line.py <report.ncu-rep> <kernel regex> [launch-id] [--top N]
The .so must be the build the report was taken from (same SASS addresses)."""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def line_table(kernel_substr):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "rustlight_b200", "librl_b200.so")], cwd=tmp, check=True, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], cwd=tmp, capture_output=True, text=True).stdout
    table, cur, infn, inl = {}, None, False, ""
    for ln in dis.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+),", ln)
        if m:
            infn = bool(re.search(kernel_substr, m.group(1)))
            continue
        if not infn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            inl = m.group(3)
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            table[int(m.group(1), 16)] = (cur, m.group(2).strip())
    return table


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = 45
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern] +
                         (["--launch-skip", sys.argv[3], "--launch-count", "1"] if len(sys.argv) > 3 and sys.argv[3].isdigit() else []),
                         capture_output=True, text=True).stdout
    # several kernels may follow each other; take the first block
    blocks = out.split('"Kernel Name"')
    blk = blocks[1]
    rows = list(csv.reader(io.StringIO(blk[blk.index("\n") + 1:])))
    hdr = rows[0]
    ia, ii, it = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
    isamp = hdr.index("# Samples")
    mangled = {"k_trace": "k_traceILb1", "k_shade": "k_shadeILb0", "k_shadow": "k_shadowILb1"}.get(kern, kern)
    table = line_table(mangled)
    base = None
    per_line = collections.Counter()
    per_line_thr = collections.Counter()
    per_line_samp = collections.Counter()
    per_op = collections.Counter()
    total = 0
    for r in rows[1:]:
        if len(r) <= it or not r[ia]:
            continue
        addr = int(r[ia], 16)
        if base is None:
            base = addr
        off = addr - base
        n = int(r[ii] or 0)
        total += n
        key, sass = table.get(off, (None, "?"))
        per_line[key] += n
        per_line_thr[key] += int(r[it] or 0)
        per_line_samp[key] += int(r[isamp] or 0)
        op = re.sub(r"^@!?U?P\d+\s+", "", sass).split(" ")[0].split(".")[0]
        per_op[op] += n
    tot_s = sum(per_line_samp.values())
    print(f"total warp instructions {total}, samples {tot_s}")
    for key, n in per_line.most_common(top):
        print(f"{100.0 * n / total:6.2f}%  inst  {100.0 * per_line_samp[key] / max(tot_s, 1):6.2f}% samp  lanes {per_line_thr[key] / max(n, 1):5.1f}  {key}")
    print("by opcode:", [(o, round(100.0 * n / total, 1)) for o, n in per_op.most_common(22)])


if __name__ == "__main__":
    main()
