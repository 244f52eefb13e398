python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s28_smoke.log 2>&1; tail -2 gpurun_out/s28_smoke.log
python bench.py --sweep --batch-scan > gpurun_out/s28_bench.json 2> gpurun_out/s28_bench.err; tail -c 600 gpurun_out/s28_bench.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/s28_ref.json 2> gpurun_out/s28_ref.err; cat gpurun_out/s28_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s28_launches.csv python bench.py --steps 12 --warmup 5 --no-cpu-baseline > gpurun_out/s28_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 10 -c 1 -o gpurun_out/s28_step_E1024 -f python bench.py --steps 12 --warmup 5 --no-cpu-baseline > gpurun_out/s28_ncu2.log 2>&1
ls -la gpurun_out/
python bench.py --tower-height 4 --max-steps 15 --no-cpu-baseline --steps 500 > gpurun_out/s28_bench_h4.json 2> gpurun_out/s28_bench_h4.err; python - <<PY
import json
d = json.load(open("gpurun_out/s28_bench_h4.json"))
print("h4", d["value"], d["ms_per_step"], d["e2e"]["value"], d["with_candidate_stage"]["value"], d["env_stats"])
PY
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
