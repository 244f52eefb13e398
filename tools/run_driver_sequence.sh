#!/bin/bash
# the driver's sequence (T=tag of the output files) -- GPU tests, smoke, default bench line (north star), reference arm
mkdir -p gpurun_out
export T=${T:-r2e}
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/${T:-r2e}_drv_pytest.log 2>&1
tail -6 gpurun_out/${T:-r2e}_drv_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T:-r2e}_drv_smoke.log 2>&1; tail -2 gpurun_out/${T:-r2e}_drv_smoke.log
( time python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/${T:-r2e}_drv_bench.json 2> gpurun_out/${T:-r2e}_drv_bench.err
tail -3 gpurun_out/${T:-r2e}_drv_bench.err
python - <<'PY'
import json, os
try:
    d = json.loads(open('gpurun_out/%s_drv_bench.json' % os.environ.get('T', 'r2e')).read().strip().splitlines()[0])
    def show(tag, r):
        ss = r.get('steady_state', {})
        print(tag, 'value %.3f M  ms %.4f | steady mean %.4f med %.4f p99 %.4f -> %.3f M | e2e %.3f M | cand %.4f ms | hbm frac %.4f fp64 frac %.5f' % (
            r['value']/1e6, r['ms_per_step'], ss.get('mean_ms', 0), ss.get('median_ms', 0), ss.get('p99_ms', 0), ss.get('value', 0)/1e6,
            r['e2e']['value']/1e6, r['with_candidate_stage']['candidate_ms_per_step'], r['roofline']['frac'], r['roofline_fp64']['frac']))
    show('headline', d)
    for k, r in d.get('secondary', {}).items():
        if isinstance(r, dict): show(k, r)
    print('e2e variants', {k: round(v['value']/1e6, 3) for k, v in d['e2e'].items() if isinstance(v, dict)})
    print('sweep', d.get('sweep', {}).get('ms_per_pass'), d.get('sweep', {}).get('value'))
    print('rollout', {k: d['rollout'][k] for k in ('value', 'ms_per_iteration', 'valid_frac', 'rasters_consistent', 'candidate_overflow')})
    print('parity_gate', d.get('parity_gate'))
    print('cpu', d.get('cpu_baseline', {}).get('value'), d.get('cpu_baseline', {}).get('cores'))
except Exception as ex:
    print('ERR', ex)
PY
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/${T:-r2e}_drv_ref.json 2> gpurun_out/${T:-r2e}_drv_ref.err; cut -c1-200 gpurun_out/${T:-r2e}_drv_ref.json
