#!/bin/bash
# candidate stage A/B: parity tests of the candidate / rollout kernels, then the candidate stage and the fused rollout
# on the bridge and tower-2 workloads (store slots; BW_BENCH_CAND_BITS=dense for the dense copies)
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_actions.py tests/test_rollout.py tests/test_gpu_rollout_parity.py tests/test_reference_rollout.py -m gpu -x -q ) > gpurun_out/c_pytest.log 2>&1
tail -15 gpurun_out/c_pytest.log
for wl in bridge tower2; do python bench.py --workload $wl --steps 200 --warmup 20 --no-cpu-baseline --no-parity-gate --steady-seconds 0.5 > gpurun_out/c_$wl.json 2>gpurun_out/c_$wl.err; done
BW_BENCH_CAND_BITS=dense python bench.py --workload bridge --steps 200 --warmup 20 --no-cpu-baseline --no-parity-gate --no-rollout --steady-seconds 0.5 > gpurun_out/c_bridge_dense.json 2>gpurun_out/c_bridge_dense.err
python - <<'PY'
import json
for wl in ("bridge","tower2","bridge_dense"):
    try:
        d=json.loads(open(f'gpurun_out/c_{wl}.json').read().strip().splitlines()[0]); ss=d['steady_state']
        r=d.get('rollout',{})
        print(wl,'value %.3f M | steady mean %.4f | cand %.4f ms (%s) | rollout %s M %s ms consistent %s'%(d['value']/1e6,ss['mean_ms'],d['with_candidate_stage']['candidate_ms_per_step'],d['with_candidate_stage'].get('candidate_rasters'),r.get('value',0)/1e6,r.get('ms_per_iteration'),r.get('rasters_consistent')))
    except Exception as ex: print(wl,'ERR',ex); print(open(f'gpurun_out/c_{wl}.err').read()[-1500:])
PY
