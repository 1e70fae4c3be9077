set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -3 > gpurun_out/r02z_multitest.log
for n in 1 2 4 8; do
  if [ $n = 1 ]; then python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02z_C2_n1.json 2> gpurun_out/r02z_C2_n1.err
  else $TR --nproc-per-node $n --master-port 2950$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r02z_C2_n$n.json 2> gpurun_out/r02z_C2_n$n.err; fi
done
for c in C4 C5 C1 C3; do
  $TR --nproc-per-node 8 --master-port 29611 bench.py --config $c --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02z_${c}_n8.json 2> gpurun_out/r02z_${c}_n8.err
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/r02z_*.json")):
    try:
        j=json.load(open(f)); print(f, j["n_gpus"], round(j["value"],1), round(j["ms_per_step"],3), "e2e", round(j["e2e"]["value"],1), j["rank_ms_per_step"], "launches", j["gpu_launches"])
    except Exception as e: print(f, "ERR", e)
PY
cat gpurun_out/r02z_multitest.log
