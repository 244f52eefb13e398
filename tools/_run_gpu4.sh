TAG=${1:-s5}
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/${TAG}_pytest.log; cat gpurun_out/${TAG}_pytest.log
L=$PWD/bridges-with-reinforcement-learning_b200/libbridges_b200_prof.so
BRIDGES_B200_LIB=$L python tools/tail_profile.py > gpurun_out/${TAG}_tail.txt 2>&1; head -3 gpurun_out/${TAG}_tail.txt; tail -14 gpurun_out/${TAG}_tail.txt
python bench.py --no-cpu-baseline --sweep --batch-scan > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_bench.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["with_candidate_stage"]["value"], d["env_stats"])
print(d["sweep"]["ms_per_pass"], [r["env_steps_per_s"] for r in d["batch_scan"]["rows"]])
PY
