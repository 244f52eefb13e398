#!/bin/bash
# round 2 (final build, LP verdict path): phase cycles, LP counters, ncu launch list, --set full captures
mkdir -p gpurun_out
PROF=$PWD/bridges-with-reinforcement-learning_b200/libbridges_b200_prof.so
T=r2b
BRIDGES_B200_LIB=$PROF timeout 300 python tools/sweep_profile.py 65536 > gpurun_out/${T}_sweep_phases.txt 2>&1
for c in bridge tower2 tower4; do
  BRIDGES_B200_LIB=$PROF timeout 300 python tools/lp_profile.py 1024 $c > gpurun_out/${T}_lp_profile_$c.txt 2>&1
  timeout 300 python tools/lp_stats.py 1024 $c 120 > gpurun_out/${T}_lp_stats_$c.txt 2>&1
done
CMD="python bench.py --steps 12 --warmup 5 --no-cpu-baseline --no-parity-gate --steady-seconds 0 --e2e-steps 12"
$CMD > gpurun_out/${T}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${T}_launches.csv $CMD > gpurun_out/${T}_ncu0.log 2>&1
CMD1="python bench.py --workload bridge --steps 12 --warmup 5 --no-cpu-baseline --no-parity-gate --no-rollout --steady-seconds 0 --e2e-steps 12"
$CMD1 > gpurun_out/${T}_plain1.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 60 -c 3 -o gpurun_out/${T}_step_bridge_E1024 -f $CMD1 > gpurun_out/${T}_ncu1.log 2>&1
CMD2="python bench.py --workload tower2 --steps 12 --warmup 5 --no-cpu-baseline --no-parity-gate --no-rollout --steady-seconds 0 --e2e-steps 12"
$CMD2 > gpurun_out/${T}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 40 -c 3 -o gpurun_out/${T}_step_tower2_E1024 -f $CMD2 > gpurun_out/${T}_ncu3.log 2>&1
CMD3="python tools/sweep_profile.py 65536 noprof"
$CMD3 > gpurun_out/${T}_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 17 -c 1 -o gpurun_out/${T}_step_sweep_E65536 -f $CMD3 > gpurun_out/${T}_ncu4.log 2>&1
ls -la gpurun_out/${T}_*.ncu-rep
