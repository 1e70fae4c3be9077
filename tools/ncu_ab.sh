for v in ref_on; do
  RL_B200_LIB=build/variants/librl_$v.so ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02k_launches_$v.csv python tools/profile_step.py 32 > /dev/null 2>&1
done
python - <<'PY'
import csv, collections
for v in ["ref_on"]:
    rows=list(csv.reader(open(f"gpurun_out/r02k_launches_{v}.csv")))
    hi=next(i for i,r in enumerate(rows) if r and r[0]=="ID")
    hdr=rows[hi]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); ui=hdr.index("Metric Unit")
    print(v)
    for r in rows[hi+1:hi+40]:
        if len(r)>vi: print("   ", r[ki][:40], r[vi], r[ui])
PY
