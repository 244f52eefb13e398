#!/bin/bash
# quick A/B numbers: three env-loop workloads (steady state) + the 65,536-assembly sweep
mkdir -p gpurun_out
for wl in bridge tower2 tower4; do python bench.py --workload $wl --steps 200 --warmup 20 --no-cpu-baseline --no-parity-gate --no-rollout --steady-seconds 1.0 $( [ $wl = bridge ] && echo --sweep ) > gpurun_out/q_$wl.json 2>gpurun_out/q_$wl.err; done
python - <<'PY'
import json
for wl in ("bridge","tower2","tower4"):
    try:
        d=json.loads(open(f'gpurun_out/q_{wl}.json').read().strip().splitlines()[0]); ss=d['steady_state']
        print(wl,'value %.3f M | steady mean %.4f med %.4f p99 %.4f | e2e %.3f M'%(d['value']/1e6,ss['mean_ms'],ss['median_ms'],ss['p99_ms'],d['e2e']['value']/1e6), ('| sweep ms %.3f'%d['sweep']['ms_per_pass']) if 'sweep' in d else '')
    except Exception as ex: print(wl,'ERR',ex)
PY
