#!/bin/bash
# round 2, job 6: profiles -- phase cycles of the sweep, ncu launch list of the default bench command, --set full
# captures of the step kernel (headline workload), the sweep evaluation and the enumerate kernel
mkdir -p gpurun_out
PROF=$PWD/bridges-with-reinforcement-learning_b200/libbridges_b200_prof.so
BRIDGES_B200_LIB=$PROF timeout 300 python tools/sweep_profile.py 65536 > gpurun_out/r2j6_sweep_phases.txt 2>&1
BRIDGES_B200_LIB=$PROF timeout 300 python tools/tail_profile.py 1024 bridge > gpurun_out/r2j6_tail_bridge.txt 2>&1
CMD="python bench.py --steps 12 --warmup 5 --no-cpu-baseline --no-parity-gate --steady-seconds 0 --e2e-steps 12"
$CMD > gpurun_out/r2j6_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/r2j6_ncu0.log 2>&1
CMD1="python bench.py --workload bridge --steps 12 --warmup 5 --no-cpu-baseline --no-parity-gate --no-rollout --steady-seconds 0 --e2e-steps 12"
$CMD1 > gpurun_out/r2j6_plain1.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 60 -c 3 -o gpurun_out/r2_step_bridge_E1024 $CMD1 > gpurun_out/r2j6_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:enumerate_kernel -s 60 -c 2 -o gpurun_out/r2_enum_bridge_E1024 $CMD1 > gpurun_out/r2j6_ncu2.log 2>&1
CMD2="python bench.py --workload tower2 --steps 12 --warmup 5 --no-cpu-baseline --no-parity-gate --no-rollout --steady-seconds 0 --e2e-steps 12"
$CMD2 > gpurun_out/r2j6_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 40 -c 3 -o gpurun_out/r2_step_tower2_E1024 $CMD2 > gpurun_out/r2j6_ncu3.log 2>&1
CMD3="python tools/sweep_profile.py 65536"
$CMD3 > gpurun_out/r2j6_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 15 -c 1 -o gpurun_out/r2_step_sweep_E65536 $CMD3 > gpurun_out/r2j6_ncu4.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
head -30 gpurun_out/r2j6_sweep_phases.txt
