#!/bin/bash
# round 2, job 1: driver-style GPU tests + tail profiles of the step kernel (BW_PROFILE build)
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2j1_pytest.log 2>&1
PROF=$PWD/bridges-with-reinforcement-learning_b200/libbridges_b200_prof.so
for c in tower2 bridge tower4; do
  BRIDGES_B200_LIB=$PROF timeout 300 python tools/tail_profile.py 1024 $c > gpurun_out/r2j1_tail_$c.txt 2>&1
done
tail -5 gpurun_out/r2j1_pytest.log
