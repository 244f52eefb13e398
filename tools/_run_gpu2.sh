python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/s3_pytest4.log; cat gpurun_out/s3_pytest4.log
L=$PWD/bridges-with-reinforcement-learning_b200/libbridges_b200_prof.so
BRIDGES_B200_LIB=$L python tools/phase_profile.py 1024 > gpurun_out/s3_phase4.txt 2>&1; tail -8 gpurun_out/s3_phase4.txt
BRIDGES_B200_LIB=$L python tools/tail_profile.py > gpurun_out/s3_tail4.txt 2>&1; head -14 gpurun_out/s3_tail4.txt
python bench.py --no-cpu-baseline --sweep --batch-scan > gpurun_out/s3_bench3.json 2> gpurun_out/s3_bench3.err; python - <<'PY'
import json
d = json.load(open("gpurun_out/s3_bench3.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["with_candidate_stage"]["value"], d["env_stats"])
print(d["sweep"]["ms_per_pass"], [r["env_steps_per_s"] for r in d["batch_scan"]["rows"]])
PY
