#!/bin/bash
# ncu --set full capture of the candidate kernel (store slots) in the steady state of the bridge workload
mkdir -p gpurun_out
T=${1:-r2c}
CMD1="python bench.py --workload bridge --steps 12 --warmup 5 --no-cpu-baseline --no-parity-gate --no-rollout --steady-seconds 0 --e2e-steps 12"
$CMD1 > gpurun_out/${T}_plain1.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:enumerate_store -s 60 -c 3 -o gpurun_out/${T}_enum_bridge_E1024 -f $CMD1 > gpurun_out/${T}_ncu_enum.log 2>&1
ls -la gpurun_out/${T}_*.ncu-rep; tail -3 gpurun_out/${T}_ncu_enum.log
# the instantiation that closes a rollout iteration (record, restart, candidates, finalize, next pick), steady state
CMD2="python tools/pipelined_rollout.py 1024 1"
$CMD2 > gpurun_out/${T}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:enumerate_store -s 100 -c 3 -o gpurun_out/${T}_enum_fin_rollout_E1024 -f $CMD2 > gpurun_out/${T}_ncu_enum2.log 2>&1
ls -la gpurun_out/${T}_*.ncu-rep; tail -2 gpurun_out/${T}_plain2.log
