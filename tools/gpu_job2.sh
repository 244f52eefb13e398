#!/bin/bash
# round 2, job 2: parity suite + A/B of the warm start (BW_NO_WARM=1 = every solve from y = 0)
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2j2_pytest.log 2>&1
tail -3 gpurun_out/r2j2_pytest.log
B="python bench.py --steps 400 --warmup 30 --no-cpu-baseline --e2e-steps 50"
for rep in 1 2; do
  BW_NO_WARM=1 $B > gpurun_out/r2j2_tower2_cold_$rep.json 2> gpurun_out/r2j2_err.txt
  $B > gpurun_out/r2j2_tower2_warm_$rep.json 2>> gpurun_out/r2j2_err.txt
done
BW_NO_WARM=1 $B --task bridge --max-steps 15 > gpurun_out/r2j2_bridge_cold.json 2>> gpurun_out/r2j2_err.txt
$B --task bridge --max-steps 15 > gpurun_out/r2j2_bridge_warm.json 2>> gpurun_out/r2j2_err.txt
BW_NO_WARM=1 $B --tower-height 4 --max-steps 15 > gpurun_out/r2j2_tower4_cold.json 2>> gpurun_out/r2j2_err.txt
$B --tower-height 4 --max-steps 15 > gpurun_out/r2j2_tower4_warm.json 2>> gpurun_out/r2j2_err.txt
PROF=$PWD/bridges-with-reinforcement-learning_b200/libbridges_b200_prof.so
for c in tower2 bridge; do
  BRIDGES_B200_LIB=$PROF timeout 300 python tools/tail_profile.py 1024 $c > gpurun_out/r2j2_tail_$c.txt 2>&1
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2j2_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']/1e6,3), 'M/s', round(d['ms_per_step'],4), 'ms  e2e', round(d['e2e']['value']/1e6,3), 'iters', round(d['env_stats']['mean_newton_iters_per_step'],2), 'cand', round(d['with_candidate_stage']['candidate_ms_per_step'],4))
    except Exception as ex:
        print(f, 'ERR', ex)
PY
