python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --sweep --batch-scan > gpurun_out/s25_bench.json 2> gpurun_out/s25_bench.err; python - <<PY
import json
d = json.load(open("gpurun_out/s25_bench.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["with_candidate_stage"]["value"], d["cpu_baseline"]["value"], d["cpu_baseline"]["one_core"], d["roofline"]["frac"], d["gpu_launches"])
PY
python bench.py --impl reference --steps 20 --warmup 3 | cut -c1-220
