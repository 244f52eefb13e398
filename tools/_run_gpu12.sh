python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 8 --steps 600 --warmup 20 --no-cpu-baseline > gpurun_out/s27_n8.json 2> gpurun_out/s27_n8.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/s27_n8.json") if l.startswith("{")][-1])
print(d["n_gpus"], d["value"], d["ms_per_step"], {k:(v["value"] if isinstance(v,dict) else v) for k,v in d["e2e"].items() if k in ("value","staged_copies","with_f32_images","with_bit_rasters")})
PY
