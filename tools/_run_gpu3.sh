python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s35_smoke.log 2>&1; tail -2 gpurun_out/s35_smoke.log
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --sweep --batch-scan > gpurun_out/s35_bench.json 2> gpurun_out/s35_bench.err; tail -c 300 gpurun_out/s35_bench.json; tail -2 gpurun_out/s35_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/s35_ref.json 2> gpurun_out/s35_ref.err; cut -c1-200 gpurun_out/s35_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s35_launches.csv python bench.py --steps 12 --warmup 5 --no-cpu-baseline > gpurun_out/s35_ncu1.log 2>&1
python bench.py --tower-height 4 --max-steps 15 --no-cpu-baseline --steps 500 > gpurun_out/s35_bench_h4.json 2> gpurun_out/s35_bench_h4.err
python examples/train_tower.py --tower-height 4 --max-steps 15 --log gpurun_out/s35_train_h4.jsonl > gpurun_out/s35_train_h4.log 2>&1; tail -1 gpurun_out/s35_train_h4.jsonl | cut -c1-300
python - <<PY
import json
for f in ("s35_bench", "s35_bench_h4"):
    d = json.load(open(f"gpurun_out/{f}.json"))
    print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["with_candidate_stage"]["value"], d["with_candidate_stage"]["candidate_ms_per_step"])
PY
