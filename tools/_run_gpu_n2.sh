python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1000 --warmup 10 > gpurun_out/s36_bench_n2.json 2> gpurun_out/s36_bench_n2.err; tail -c 400 gpurun_out/s36_bench_n2.json; tail -2 gpurun_out/s36_bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/s36_ref_n2.json 2> gpurun_out/s36_ref_n2.err; cut -c1-160 gpurun_out/s36_ref_n2.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --task bridge --max-steps 15 --steps 300 --warmup 10 > gpurun_out/s36_bridge_n2.json 2> gpurun_out/s36_bridge_n2.err
python - <<PY
import json
for f in ("s36_bench_n2", "s36_bridge_n2"):
    d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    print(f, d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["with_candidate_stage"]["value"])
PY
