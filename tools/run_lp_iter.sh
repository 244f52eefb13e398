#!/bin/bash
# iteration job of the LP path: quick parity subset, LP profile (prof build), counters, A/B-less bench of two workloads
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_step.py tests/test_gpu_rollout_parity.py tests/test_gpu_dropin.py -x -q ) > gpurun_out/lpi_pytest.log 2>&1; tail -3 gpurun_out/lpi_pytest.log
PROF=bridges-with-reinforcement-learning_b200/libbridges_b200_prof.so
if [ -f $PROF ]; then for c in bridge tower2; do BRIDGES_B200_LIB=$PROF python tools/lp_profile.py 1024 $c > gpurun_out/lp_profile_$c.txt 2>&1; done; fi
python tools/lp_stats.py 1024 bridge 120 > gpurun_out/lp_stats_bridge.txt 2>&1
for wl in bridge tower2 tower4; do python bench.py --workload $wl --steps 200 --warmup 20 --no-cpu-baseline --no-parity-gate --no-rollout --steady-seconds 1.0 > gpurun_out/lp_bench_$wl.json 2>gpurun_out/lp_bench_$wl.err; done
python - <<'PY'
import json
for wl in ("bridge","tower2","tower4"):
    try:
        d=json.loads(open(f'gpurun_out/lp_bench_{wl}.json').read().strip().splitlines()[0])
        ss=d['steady_state']
        print(wl,'value %.3f M ms %.4f | steady mean %.4f med %.4f p99 %.4f | e2e %.3f M | cand %.4f ms'%(d['value']/1e6,d['ms_per_step'],ss['mean_ms'],ss['median_ms'],ss['p99_ms'],d['e2e']['value']/1e6,d['with_candidate_stage']['candidate_ms_per_step']))
    except Exception as ex: print(wl,'ERR',ex)
PY
