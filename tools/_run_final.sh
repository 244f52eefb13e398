python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s38_smoke.log 2>&1; tail -2 gpurun_out/s38_smoke.log
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/s38_bench.json 2> gpurun_out/s38_bench.err; tail -c 300 gpurun_out/s38_bench.json; tail -2 gpurun_out/s38_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s38_launches.csv python bench.py --steps 12 --warmup 5 --no-cpu-baseline > gpurun_out/s38_ncu1.log 2>&1
python bench.py --task bridge --max-steps 15 --steps 500 --no-cpu-baseline > gpurun_out/s38_bench_bridge.json 2> gpurun_out/s38_bench_bridge.err
python - <<PY
import json
for f in ("s38_bench", "s38_bench_bridge"):
    d = json.load(open(f"gpurun_out/{f}.json"))
    print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["with_candidate_stage"]["value"], d["with_candidate_stage"]["candidate_ms_per_step"], d["gpu_launches"])
PY
