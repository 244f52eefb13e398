#!/bin/bash
# proximal-schedule experiment: kernel time, Newton steps and verdict parity per schedule
for S in "1e2,1e4,1e6,1e8" "1e3,1e6,1e9" "1e2,1e5,1e8" "1e4,1e8" "1e3,1e5,1e7,1e9" "1e1,1e3,1e5,1e7,1e9,1e9"; do
  echo "=== $S"
  BW_RHO_SCHEDULE=$S python bench.py --no-cpu-baseline --sweep --steps 500 --e2e-steps 10 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms_per_step %.4f iters %.2f notconv %d | sweep ms %.2f iters %.2f notconv %d stable %.4f su %.4f' % (d['ms_per_step'], d['env_stats']['mean_newton_iters_per_step'], d['env_stats']['solver_not_converged'], d['sweep']['ms_per_pass'], d['sweep']['rank0_stats']['mean_newton_iters'], d['sweep']['rank0_stats']['not_converged'], d['sweep']['rank0_stats']['stable_frac'], d['sweep']['rank0_stats']['stable_unfrozen_frac']))"
  BW_RHO_SCHEDULE=$S python -m pytest tests/test_gpu_step.py -m gpu -x -q 2>&1 | tail -2
done
