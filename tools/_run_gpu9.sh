timeout 500 compute-sanitizer --tool memcheck --print-limit 5 python tools/sanitize_smoke.py > gpurun_out/s22_memcheck.log 2>&1; tail -6 gpurun_out/s22_memcheck.log
timeout 500 compute-sanitizer --tool racecheck --print-limit 5 python tools/sanitize_smoke.py > gpurun_out/s22_racecheck.log 2>&1; tail -6 gpurun_out/s22_racecheck.log
