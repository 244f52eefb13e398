#!/bin/bash
# one --set full capture of step_kernel on a workload (after the same command ran clean without ncu)
# usage: tools/run_ncu_step.sh <workload> <tag> [skip]
WL=${1:-bridge}; TAG=${2:-r2lp}; SKIP=${3:-60}
mkdir -p gpurun_out
CMD="python bench.py --workload $WL --steps 12 --warmup 5 --no-cpu-baseline --no-parity-gate --no-rollout --steady-seconds 0 --e2e-steps 12"
$CMD > gpurun_out/${TAG}_plain_$WL.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s $SKIP -c 3 -o gpurun_out/${TAG}_step_${WL}_E1024 -f $CMD > gpurun_out/${TAG}_ncu_$WL.log 2>&1
ls -la gpurun_out/${TAG}_step_${WL}_E1024.ncu-rep
