ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:step_kernelILb1 -s 20 -c 1 -o gpurun_out/s23_sweep_E65536 -f python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 2 --sweep > gpurun_out/s23_ncu.log 2>&1
tail -2 gpurun_out/s23_ncu.log
