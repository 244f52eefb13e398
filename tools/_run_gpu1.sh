python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/s3_pytest3.log; cat gpurun_out/s3_pytest3.log
L=$PWD/bridges-with-reinforcement-learning_b200/libbridges_b200_prof.so
BRIDGES_B200_LIB=$L python tools/phase_profile.py 128 > gpurun_out/s3_phase_E128.txt 2>&1
BRIDGES_B200_LIB=$L python tools/phase_profile.py 1024 > gpurun_out/s3_phase_E1024.txt 2>&1
BRIDGES_B200_LIB=$L python tools/phase_profile.py 8192 > gpurun_out/s3_phase_E8192.txt 2>&1
for f in gpurun_out/s3_phase_E128.txt gpurun_out/s3_phase_E1024.txt gpurun_out/s3_phase_E8192.txt; do echo == $f; head -12 $f; done
python bench.py --no-cpu-baseline --sweep --batch-scan > gpurun_out/s3_bench2.json 2> gpurun_out/s3_bench2.err; python - <<'PY'
import json
d = json.load(open("gpurun_out/s3_bench2.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["with_candidate_stage"]["value"], d["env_stats"])
print(d["sweep"]); print(d["batch_scan"])
PY
