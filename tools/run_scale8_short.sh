set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -3 > gpurun_out/r03_multitest.log
$TR --nproc-per-node 8 --master-port 29508 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r03_C2_n8.json 2> gpurun_out/r03_C2_n8.err
$TR --nproc-per-node 8 --master-port 29611 bench.py --config C4 --gpus 8 --steps 5 --warmup 3 > gpurun_out/r03_C4_n8.json 2> gpurun_out/r03_C4_n8.err
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/r03_C*_n8.json")):
    try:
        j=json.load(open(f)); print(f, j["n_gpus"], round(j["value"],1), round(j["ms_per_step"],3), "e2e", round(j["e2e"]["value"],1), j["rank_ms_per_step"], "launches", j["gpu_launches"])
    except Exception as e: print(f, "ERR", e)
PY
cat gpurun_out/r03_multitest.log
