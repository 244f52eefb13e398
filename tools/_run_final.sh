python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s41_smoke.log 2>&1; tail -2 gpurun_out/s41_smoke.log
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --no-cpu-baseline --steps 1000 > gpurun_out/s41_bench.json 2> gpurun_out/s41_bench.err
python bench.py --task bridge --max-steps 15 --steps 500 --no-cpu-baseline > gpurun_out/s41_bench_bridge.json 2> gpurun_out/s41_bench_bridge.err
python - <<PY
import json
for f in ("s41_bench", "s41_bench_bridge"):
    d = json.load(open(f"gpurun_out/{f}.json"))
    print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["with_candidate_stage"]["value"], d["with_candidate_stage"]["candidate_ms_per_step"], d["gpu_launches"])
PY
