#!/bin/bash
# Collects everything profiles/ summarises (run under gpurun on one B200): tools/collect_profiles.sh <tag>
tag=${1:-r02}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/${tag}_gputest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench_C2.json 2> gpurun_out/${tag}_bench_C2.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${tag}_bench_C2_ref.json 2>> gpurun_out/${tag}_bench_C2.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_trace_shadow_flat -c 3 -f -o gpurun_out/${tag}_trace python tools/profile_step.py 8 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_shade -c 3 -f -o gpurun_out/${tag}_shade python tools/profile_step.py 8 > /dev/null 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_trace_shadow_flat --csv --log-file gpurun_out/${tag}_traffic_trace.csv python tools/profile_step.py 128 > /dev/null 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_shade --csv --log-file gpurun_out/${tag}_traffic_shade.csv python tools/profile_step.py 128 > /dev/null 2>&1
cat gpurun_out/${tag}_gputest.log; cut -c1-400 gpurun_out/${tag}_bench_C2.json; ls -la gpurun_out/${tag}_*
