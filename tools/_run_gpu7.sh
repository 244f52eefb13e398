nvidia-smi -L | head -8
for N in 1 2 4 8; do
  if [ $N = 1 ]; then
    python bench.py --gpus 1 --steps 1000 --warmup 20 --no-cpu-baseline > gpurun_out/s16_scale_n1.json 2> gpurun_out/s16_scale_n1.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 1000 --warmup 20 --no-cpu-baseline > gpurun_out/s16_scale_n$N.json 2> gpurun_out/s16_scale_n$N.err
  fi
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/s16_scale_n$N.json") if l.startswith("{")][-1])
    print($N, d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"])
except Exception as ex:
    print("N=$N failed", ex)
PY
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tools/multi_gpu_rollout.py > gpurun_out/s16_rollout_n8.json 2> gpurun_out/s16_rollout_n8.err; tail -3 gpurun_out/s16_rollout_n8.json
