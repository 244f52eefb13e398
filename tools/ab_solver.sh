#!/bin/bash
# solver A/B on the GPU: kernel time, Newton steps and verdict parity per (schedule, screen) setting
# usage: tools/ab_solver.sh OUTFILE  ["SCHED;NOSCREEN" ...]
OUT=$1; shift
for V in "$@"; do
  S=${V%%;*}; NS=${V##*;}
  echo "=== schedule=$S no_screen=$NS" >> $OUT
  ( [ "$S" != "default" ] && export BW_RHO_SCHEDULE=$S; [ "$NS" = "1" ] && export BW_NO_SCREEN=1
    python bench.py --no-cpu-baseline --sweep --batch-scan --steps 500 --e2e-steps 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
bs=d.get('batch_scan', {})
print('ms_per_step %.4f value %.0f iters %.2f notconv %d | sweep ms %.2f iters %.2f notconv %d stable %.4f su %.4f | scan %s' % (d['ms_per_step'], d['value'], d['env_stats']['mean_newton_iters_per_step'], d['env_stats']['solver_not_converged'], d['sweep']['ms_per_pass'], d['sweep']['rank0_stats']['mean_newton_iters'], d['sweep']['rank0_stats']['not_converged'], d['sweep']['rank0_stats']['stable_frac'], d['sweep']['rank0_stats']['stable_unfrozen_frac'], json.dumps(bs)[:400]))" >> $OUT 2>&1
    python -m pytest tests/test_gpu_step.py tests/test_gpu_properties.py -m gpu -x -q 2>&1 | tail -2 >> $OUT )
done
