ncu --set full --clock-control none --import-source on -k regex:enumerate_kernel -s 10 -c 1 -o gpurun_out/s12_enum -f python bench.py --steps 12 --warmup 5 --no-cpu-baseline > gpurun_out/s12_ncu.log 2>&1
tail -2 gpurun_out/s12_ncu.log
