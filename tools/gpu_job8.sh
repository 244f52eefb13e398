#!/bin/bash
# round 2, job 8: shared packed matrix (share_h) -- parity suite in both modes, then A/B
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2j8_pytest.log 2>&1
tail -4 gpurun_out/r2j8_pytest.log
( BW_SHARE_H=2 timeout 1500 python -m pytest tests/test_gpu_step.py tests/test_gpu_rollout_parity.py tests/test_gpu_properties.py -m gpu -q -x ) > gpurun_out/r2j8_pytest_share.log 2>&1
tail -4 gpurun_out/r2j8_pytest_share.log
B="python bench.py --steps 300 --warmup 30 --no-cpu-baseline --no-parity-gate --no-rollout --e2e-steps 50 --steady-seconds 1.0"
for w in bridge tower4; do
  BW_SHARE_H=0 $B --workload $w > gpurun_out/r2j8_${w}_twoH.json 2> gpurun_out/r2j8_err.txt
  $B --workload $w > gpurun_out/r2j8_${w}_auto.json 2>> gpurun_out/r2j8_err.txt
done
BW_SHARE_H=0 $B --workload tower2 --sweep --sweep-steps 20 --batch-scan > gpurun_out/r2j8_sweep_twoH.json 2>> gpurun_out/r2j8_err.txt
$B --workload tower2 --sweep --sweep-steps 20 --batch-scan > gpurun_out/r2j8_sweep_auto.json 2>> gpurun_out/r2j8_err.txt
BW_SHARE_H=1 $B --workload tower2 --sweep --sweep-steps 20 > gpurun_out/r2j8_sweep_share1.json 2>> gpurun_out/r2j8_err.txt
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2j8_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        ss=d.get('steady_state',{})
        print(f, 'value %.3f M (%.4f ms) steady mean %.4f med %.4f p99 %.4f' % (d['value']/1e6, d['ms_per_step'], ss.get('mean_ms',0), ss.get('median_ms',0), ss.get('p99_ms',0)),
              'sweep %.3f ms' % d['sweep']['ms_per_pass'] if 'sweep' in d else '', [ (r['envs'], round(r['env_steps_per_s']/1e6,2)) for r in d.get('batch_scan',{}).get('rows',[])])
    except Exception as ex:
        print(f, 'ERR', ex)
PY
tail -3 gpurun_out/r2j8_err.txt
