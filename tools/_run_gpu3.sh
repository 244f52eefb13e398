python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s10_smoke.log 2>&1; tail -2 gpurun_out/s10_smoke.log
python bench.py --sweep --batch-scan > gpurun_out/s10_bench.json 2> gpurun_out/s10_bench.err; tail -c 600 gpurun_out/s10_bench.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/s10_ref.json 2> gpurun_out/s10_ref.err; cat gpurun_out/s10_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s10_launches.csv python bench.py --steps 12 --warmup 5 --no-cpu-baseline > gpurun_out/s10_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 10 -c 1 -o gpurun_out/s10_step_E1024 -f python bench.py --steps 12 --warmup 5 --no-cpu-baseline > gpurun_out/s10_ncu2.log 2>&1
ls -la gpurun_out/
