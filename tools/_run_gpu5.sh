TAG=${1:-s7}
L=$PWD/bridges-with-reinforcement-learning_b200/libbridges_b200_prof.so
BRIDGES_B200_LIB=$L python tools/tail_profile.py > gpurun_out/${TAG}_tail.txt 2>&1; head -3 gpurun_out/${TAG}_tail.txt; tail -14 gpurun_out/${TAG}_tail.txt
