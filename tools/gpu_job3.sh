#!/bin/bash
# round 2, job 3: full parity suite (incl. fused rollout vs oracle) + warm-start A/B (frozen solve only) + rollout rates
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r2j3_pytest.log 2>&1
tail -15 gpurun_out/r2j3_pytest.log
B="python bench.py --steps 400 --warmup 30 --no-cpu-baseline --e2e-steps 50"
BW_NO_WARM=1 $B > gpurun_out/r2j3_tower2_cold.json 2> gpurun_out/r2j3_err.txt
$B > gpurun_out/r2j3_tower2_warm.json 2>> gpurun_out/r2j3_err.txt
BW_NO_WARM=1 $B --task bridge --max-steps 15 > gpurun_out/r2j3_bridge_cold.json 2>> gpurun_out/r2j3_err.txt
$B --task bridge --max-steps 15 > gpurun_out/r2j3_bridge_warm.json 2>> gpurun_out/r2j3_err.txt
$B --tower-height 4 --max-steps 15 > gpurun_out/r2j3_tower4_warm.json 2>> gpurun_out/r2j3_err.txt
for t in tower bridge; do
  python tools/multi_gpu_rollout.py 1024 $t > gpurun_out/r2j3_rollout_${t}_E1024.json 2>> gpurun_out/r2j3_err.txt
  python tools/multi_gpu_rollout.py 4096 $t > gpurun_out/r2j3_rollout_${t}_E4096.json 2>> gpurun_out/r2j3_err.txt
done
PROF=$PWD/bridges-with-reinforcement-learning_b200/libbridges_b200_prof.so
BRIDGES_B200_LIB=$PROF timeout 300 python tools/tail_profile.py 1024 tower2 > gpurun_out/r2j3_tail_tower2.txt 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2j3_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        if 'value' in d:
            print(f, round(d['value']/1e6,3), 'M/s', round(d['ms_per_step'],4), 'ms  e2e', round(d['e2e']['value']/1e6,3), 'iters', round(d['env_stats']['mean_newton_iters_per_step'],2), 'cand', round(d['with_candidate_stage']['candidate_ms_per_step'],4))
        else:
            print(f, d)
    except Exception as ex:
        print(f, 'ERR', ex)
PY
tail -5 gpurun_out/r2j3_err.txt
