#!/bin/bash
# round 2, job 7: heaviest-first CTA order of the step kernel -- parity suite, then A/B (BW_NO_ORDER=1 = index order)
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2j7_pytest.log 2>&1
tail -4 gpurun_out/r2j7_pytest.log
B="python bench.py --steps 300 --warmup 30 --no-cpu-baseline --no-parity-gate --no-rollout --e2e-steps 50 --steady-seconds 1.0"
for w in bridge tower4 tower2; do
  BW_NO_ORDER=1 $B --workload $w > gpurun_out/r2j7_${w}_index.json 2> gpurun_out/r2j7_err.txt
  $B --workload $w > gpurun_out/r2j7_${w}_heavy.json 2>> gpurun_out/r2j7_err.txt
done
BW_NO_ORDER=1 $B --workload tower2 --sweep --sweep-steps 20 > gpurun_out/r2j7_sweep_index.json 2>> gpurun_out/r2j7_err.txt
$B --workload tower2 --sweep --sweep-steps 20 > gpurun_out/r2j7_sweep_heavy.json 2>> gpurun_out/r2j7_err.txt
PROF=$PWD/bridges-with-reinforcement-learning_b200/libbridges_b200_prof.so
BRIDGES_B200_LIB=$PROF timeout 300 python tools/sweep_profile.py 65536 > gpurun_out/r2j7_sweep_phases.txt 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2j7_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        ss=d.get('steady_state',{})
        print(f, 'value %.3f M (%.4f ms) steady mean %.4f med %.4f p99 %.4f' % (d['value']/1e6, d['ms_per_step'], ss.get('mean_ms',0), ss.get('median_ms',0), ss.get('p99_ms',0)),
              'sweep %.3f ms' % d['sweep']['ms_per_pass'] if 'sweep' in d else '')
    except Exception as ex:
        print(f, 'ERR', ex)
PY
tail -3 gpurun_out/r2j7_err.txt
head -12 gpurun_out/r2j7_sweep_phases.txt
