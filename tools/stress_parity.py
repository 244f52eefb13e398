"""One-off wider run of the random-assembly parity test (verdicts, residuals, forces, poses, rasters against
the oracle) on seeds / sizes other than the ones pinned in tests/test_gpu_step.py.  Test infrastructure.

python tools/stress_parity.py [seed0]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.test_gpu_step import test_random_assemblies_verdicts_residuals_forces as run

seed0 = int(sys.argv[1]) if len(sys.argv) > 1 else 100
for N, max_blocks, max_steps, seed in ((256, 14, None, seed0), (256, 15, None, seed0 + 1), (192, 10, 10, seed0 + 2),
                                       (192, 10, 10, seed0 + 3)):
    t0 = time.perf_counter()
    run(N, max_blocks, max_steps, seed, 1)
    print(f"ok N={N} max_blocks={max_blocks} max_steps={max_steps} seed={seed}  {time.perf_counter() - t0:.1f} s", flush=True)
