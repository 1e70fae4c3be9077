# C2 at 1 / 2 / 4 / 8 GPUs of one box (development aid; the driver runs its own scaling bench): bash tools/run_scale_c2.sh <tag>
tag=${1:-r03f}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_C2_n1.json 2> gpurun_out/${tag}_C2_n1.err
for n in 2 4 8; do
  $TR --nproc-per-node $n --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/${tag}_C2_n$n.json 2> gpurun_out/${tag}_C2_n$n.err
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${tag}_C2_n*.json")):
    try:
        j=json.load(open(f)); print(f, j["n_gpus"], round(j["value"],1), round(j["ms_per_step"],3), "e2e", round(j["e2e"]["value"],1), j["rank_ms_per_step"], "launches", j["gpu_launches"])
    except Exception as e: print(f, "ERR", e)
PY
