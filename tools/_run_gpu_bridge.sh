python bench.py --task bridge --num-obstacles 5 --shapes trapezoid,hexagon --max-steps 15 --steps 500 --cpu-budget 12 > gpurun_out/s29_bench_bridge.json 2> gpurun_out/s29_bench_bridge.err; tail -c 1500 gpurun_out/s29_bench_bridge.json; tail -3 gpurun_out/s29_bench_bridge.err
python bench.py --tower-height 4 --max-steps 15 --no-cpu-baseline --steps 500 > gpurun_out/s29_bench_h4.json 2> gpurun_out/s29_bench_h4.err; python - <<PY
import json
for f in ("s29_bench_bridge", "s29_bench_h4"):
    d = json.load(open(f"gpurun_out/{f}.json"))
    print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["with_candidate_stage"]["value"], d["env_stats"], d.get("cpu_baseline", {}).get("value"))
PY
