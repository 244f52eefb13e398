#!/bin/bash
# round 2, job 10: warp-autonomous enumerate kernel -- candidate / rollout parity tests, then A/B against the previous
# build (libbridges_b200_prev.so) on the candidate stage and the fused rollout
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2j10_pytest.log 2>&1
tail -4 gpurun_out/r2j10_pytest.log
PREV=$PWD/bridges-with-reinforcement-learning_b200/libbridges_b200_prev.so
B="python bench.py --steps 300 --warmup 30 --no-cpu-baseline --no-parity-gate --e2e-steps 50 --steady-seconds 0.5"
for w in bridge tower2 tower4; do
  BRIDGES_B200_LIB=$PREV $B --workload $w > gpurun_out/r2j10_${w}_prev.json 2> gpurun_out/r2j10_err.txt
  $B --workload $w > gpurun_out/r2j10_${w}_new.json 2>> gpurun_out/r2j10_err.txt
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2j10_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.3f M' % (d['value']/1e6), 'cand %.4f ms with_cand %.3f M' % (d['with_candidate_stage']['candidate_ms_per_step'], d['with_candidate_stage']['value']/1e6),
              'rollout %.3f M (%.4f ms/it)' % (d['rollout']['value']/1e6, d['rollout']['ms_per_iteration']))
    except Exception as ex:
        print(f, 'ERR', ex)
PY
tail -3 gpurun_out/r2j10_err.txt
