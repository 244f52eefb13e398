#!/bin/bash
# LP verdict path: smoke, GPU parity suite, A/B bench of the three env-loop workloads (BW_NO_LP=1 = the Newton path)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/lp_smoke.log 2>&1; tail -3 gpurun_out/lp_smoke.log
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/lp_pytest.log 2>&1
tail -15 gpurun_out/lp_pytest.log
for wl in bridge tower2 tower4; do
  for nolp in 0 1; do
    if [ $nolp = 1 ]; then export BW_NO_LP=1; else unset BW_NO_LP; fi
    timeout 300 python bench.py --workload $wl --steps 200 --warmup 20 --no-cpu-baseline --no-parity-gate --no-rollout --steady-seconds 1.0 \
       > gpurun_out/lp_bench_${wl}_nolp${nolp}.json 2> gpurun_out/lp_bench_${wl}_nolp${nolp}.err
    python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/lp_bench_${wl}_nolp${nolp}.json').read().strip().splitlines()[0])
    ss = d.get('steady_state', {})
    print('${wl} nolp=${nolp}: value %.3f M ms %.4f | steady mean %.4f med %.4f p99 %.4f | e2e %.3f M | env_stats %s' % (
        d['value']/1e6, d['ms_per_step'], ss.get('mean_ms', 0), ss.get('median_ms', 0), ss.get('p99_ms', 0), d['e2e']['value']/1e6, d.get('env_stats')))
except Exception as ex:
    print('ERR ${wl} ${nolp}', ex); print(open('gpurun_out/lp_bench_${wl}_nolp${nolp}.err').read()[-1500:])
PY
  done
done
