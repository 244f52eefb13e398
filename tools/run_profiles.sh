#!/bin/bash
# round 2, job 9: profiles of the final build (phase cycles, ncu launch list, --set full captures) + learner runs
mkdir -p gpurun_out
PROF=$PWD/bridges-with-reinforcement-learning_b200/libbridges_b200_prof.so
BRIDGES_B200_LIB=$PROF timeout 300 python tools/sweep_profile.py 65536 > gpurun_out/r2_sweep_phases.txt 2>&1
for c in bridge tower2 tower4; do
  BRIDGES_B200_LIB=$PROF timeout 300 python tools/tail_profile.py 1024 $c > gpurun_out/r2_tail_envs_$c.txt 2>&1
done
CMD="python bench.py --steps 12 --warmup 5 --no-cpu-baseline --no-parity-gate --steady-seconds 0 --e2e-steps 12"
$CMD > gpurun_out/r2j9_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/r2j9_ncu0.log 2>&1
CMD1="python bench.py --workload bridge --steps 12 --warmup 5 --no-cpu-baseline --no-parity-gate --no-rollout --steady-seconds 0 --e2e-steps 12"
$CMD1 > gpurun_out/r2j9_plain1.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 60 -c 3 -o gpurun_out/r2_step_bridge_E1024 -f $CMD1 > gpurun_out/r2j9_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:enumerate_kernel -s 60 -c 2 -o gpurun_out/r2_enum_bridge_E1024 -f $CMD1 > gpurun_out/r2j9_ncu2.log 2>&1
CMD2="python bench.py --workload tower2 --steps 12 --warmup 5 --no-cpu-baseline --no-parity-gate --no-rollout --steady-seconds 0 --e2e-steps 12"
$CMD2 > gpurun_out/r2j9_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 40 -c 3 -o gpurun_out/r2_step_tower2_E1024 -f $CMD2 > gpurun_out/r2j9_ncu3.log 2>&1
CMD3="python tools/sweep_profile.py 65536 noprof"
$CMD3 > gpurun_out/r2j9_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 15 -c 1 -o gpurun_out/r2_step_sweep_E65536 -f $CMD3 > gpurun_out/r2j9_ncu4.log 2>&1
ls -la gpurun_out/r2_*.ncu-rep
timeout 200 python examples/train_successor.py --tower-height 4 --max-steps 15 --envs 256 --seconds 120 --log gpurun_out/r2_train_successor_h4.jsonl > gpurun_out/r2j9_train_succ.log 2>&1
tail -3 gpurun_out/r2j9_train_succ.log | cut -c1-250
timeout 120 python examples/train_tower.py --tower-height 4 --max-steps 15 --iters 12 --log gpurun_out/r2_train_tower_h4.jsonl > gpurun_out/r2j9_train_tower.log 2>&1
tail -2 gpurun_out/r2j9_train_tower.log | cut -c1-250
