python -m pytest tests/test_gpu_actions.py tests/test_gpu_dropin.py tests/test_rollout.py -x -q 2>&1 | tail -5
python bench.py --no-cpu-baseline --steps 1000 > gpurun_out/s34_bench.json 2> gpurun_out/s34_bench.err
python bench.py --task bridge --max-steps 15 --steps 500 --no-cpu-baseline > gpurun_out/s34_bench_bridge.json 2> gpurun_out/s34_bench_bridge.err

python - <<PY
import json
for f in ("s34_bench", "s34_bench_bridge"):
    d = json.load(open(f"gpurun_out/{f}.json"))
    print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["with_candidate_stage"])
PY
