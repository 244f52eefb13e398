"""Scratch: Lawson-Hanson NNLS on the cone edge rays (oracle/nnls.py) against the numpy emulation of the CUDA
solver (tools/solver_lab.py) on harvested systems (tools/harvest_systems.py): verdict agreement, iteration counts
and a serial-depth cost model, before spending GPU time on a second solver.  Test infrastructure only.

python tools/nnls_lab.py /tmp/systems.pkl [n_systems]

Cost model (cycles of a lone warp, from profiles/r1_tail_envs_v6.txt and r1_latency_microbench.txt):
  Newton step of the current solver   6,000 + 8,400 m / 27   (+ 600 per extra line-search evaluation)
  NNLS least-squares solve            300 + 120 p            (two reductions + three sweeps of length p = passive
                                                              columns at ~40 cycles per dependent step)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import solver_lab as SL
from oracle import nnls
from oracle import stability as st

if __name__ == "__main__":
    systems = SL.load(sys.argv[1])
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
    rng = np.random.default_rng(1)
    rows = []
    for i in rng.permutation(len(systems))[:n]:
        A, b, mu, ok, nbk, tag = systems[i]
        status, r, it, ev = SL.solve(A, b, mu)
        m = A.shape[0]
        newton = it * (6000 + 8400 * m / 27.0) + max(0, ev - it) * 600
        rn, out = nnls.equilibrium_residual_nnls(A, b, mu)
        cost = 300 * out.iterations + 120 * out.chain
        v_newton = (status == 0) or (status == 2 and r <= 1e-6)
        rows.append((nbk, m, -1 if ok is None else int(ok), int(v_newton), int(rn <= 1e-6), it, out.iterations, newton, cost, rn))
    rows = np.array(rows)
    band = (rows[:, 9] > 1e-9) & (rows[:, 9] < 1e-4)
    print(f"{len(rows)} systems; verdict != HiGHS label: newton {int(((rows[:, 2] != rows[:, 3]) & ~band).sum())}, "
          f"nnls {int(((rows[:, 2] != rows[:, 4]) & ~band).sum())}; inside the residual band: {int(band.sum())}")
    for lo, hi in ((2, 3), (4, 6), (7, 16)):
        q = rows[(rows[:, 0] >= lo) & (rows[:, 0] <= hi)]
        if len(q):
            print(f"blocks {lo}-{hi}: n={len(q)} | newton steps mean {q[:, 5].mean():.1f} max {q[:, 5].max():.0f}, "
                  f"model cycles mean {q[:, 7].mean() / 1e3:.0f}k p99 {np.percentile(q[:, 7], 99) / 1e3:.0f}k max {q[:, 7].max() / 1e3:.0f}k"
                  f" | nnls solves mean {q[:, 6].mean():.1f} max {q[:, 6].max():.0f}, "
                  f"model cycles mean {q[:, 8].mean() / 1e3:.0f}k p99 {np.percentile(q[:, 8], 99) / 1e3:.0f}k max {q[:, 8].max() / 1e3:.0f}k")
