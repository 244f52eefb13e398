nproc; nvidia-smi topo -m 2>/dev/null | head -12
for MODE in bind nobind; do
  FLAG=""; [ $MODE = nobind ] && FLAG="--no-numa-bind"
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 600 --warmup 20 --no-cpu-baseline $FLAG > gpurun_out/s17_n8_$MODE.json 2> gpurun_out/s17_n8_$MODE.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/s17_n8_$MODE.json") if l.startswith("{")][-1])
    print("$MODE", d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"].get("host_binding"), d["e2e"]["staged_copies"]["value"])
except Exception as ex:
    print("$MODE failed", ex)
PY
done
