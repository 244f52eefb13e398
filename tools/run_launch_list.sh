#!/bin/bash
# ncu launch lists (gpu__time_duration.sum per launch; cold-cache and serialised: shares, not durations):
#   ${T}_launches.csv          the default bench command (first 1,500 launches)
#   ${T}_rollout_launches.csv  600 launches of the fused rollout (bw_rollout_random) in its steady state
mkdir -p gpurun_out
T=${T:-r2e}
CMD="python bench.py --steps 12 --warmup 5 --no-cpu-baseline --no-parity-gate --steady-seconds 0 --e2e-steps 12"
$CMD > gpurun_out/${T}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${T}_launches.csv $CMD > gpurun_out/${T}_ncu0.log 2>&1
CMD2="python tools/pipelined_rollout.py 1024 1"
$CMD2 > gpurun_out/${T}_plain2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 600 --csv --log-file gpurun_out/${T}_rollout_launches.csv $CMD2 > gpurun_out/${T}_ncu2.log 2>&1
tail -2 gpurun_out/${T}_plain2.log; wc -l gpurun_out/${T}_launches.csv gpurun_out/${T}_rollout_launches.csv
