ncu --set full --clock-control none --import-source on -k regex:enumerate_kernel -s 8 -c 1 -o gpurun_out/s39_enum_tower -f python bench.py --steps 12 --warmup 5 --no-cpu-baseline > gpurun_out/s39_ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:enumerate_kernel -s 8 -c 1 -o gpurun_out/s39_enum_bridge -f python bench.py --task bridge --max-steps 15 --steps 12 --warmup 5 --no-cpu-baseline > gpurun_out/s39_ncu_b.log 2>&1
ls -la gpurun_out/ | tail -3
