"""Counters of the LP verdict path on a bench-like rollout (needs a GPU; sets BW_LP_STATS).
python tools/lp_stats.py [E] [bridge|tower2|tower4] [steps]"""
import os, sys
os.environ["BW_LP_STATS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bridges_b200.envs.batched import BatchedAssemblyGym
from bench import task_def, bridge_def, X_GROUND
E = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
CASE = sys.argv[2] if len(sys.argv) > 2 else "bridge"
STEPS = int(sys.argv[3]) if len(sys.argv) > 3 else 120
if CASE == "bridge":
    env = BatchedAssemblyGym(E, ["shapes/trapezoid.urdf", "shapes/hexagon.urdf"], max_steps=15)
    env.reset(bridge_def(5)); AMAX = 1024
elif CASE == "tower4":
    env = BatchedAssemblyGym(E, ["shapes/trapezoid.urdf"], max_steps=15)
    env.reset(task_def(4)); AMAX = 256
else:
    env = BatchedAssemblyGym(E, ["shapes/trapezoid.urdf"], max_steps=10)
    env.reset(task_def(2)); AMAX = 128
rows = []
for i in range(STEPS):
    env.enumerate_actions(X_GROUND, (0.0,), amax=AMAX, with_bits=False)
    acts, _ = env.select_random(seed=i)
    env.step(acts)
    rows.append(env.read_out().copy())
    env.reset_done()
st = env.lp_stats()
o = np.concatenate(rows)
names = {0: "runs frozen", 1: "runs released", 2: "feasible frozen", 3: "feasible released", 4: "infeasible frozen",
         5: "infeasible released", 6: "not certified frozen", 7: "not certified released", 9: "why: pivot cap",
         10: "why: no pivot row", 11: "why: primal residual", 12: "why: dual certificate (ray)", 13: "why: dual certificate (objective)",
         14: "why: setup", 16: "pivots", 17: "max pivots of a run", 20: "rows of all runs"}
print("case", CASE, "E", E, "steps", STEPS, "env steps", len(o))
for k in sorted(names):
    print("  %-36s %d" % (names[k], st[k]))
runs = max(int(st[0] + st[1]), 1)
print("  pivots per run %.2f, rows per run %.1f, pivots without progress %.1f %%" % (st[16] / runs, st[20] / runs, 100.0 * st[23] / max(int(st[16]), 1)))
if st[24]:
    print("  runs of >= 20 pivots: %d (frozen %d, released %d; feasible %d, infeasible %d), %.1f pivots each of which %.1f without progress, %.1f rows" % (st[24], st[28], st[29], st[30], st[31], st[25] / st[24], st[26] / st[24], st[27] / st[24]))
piv = o["lp_pivots"]
print("lp_pivots per step: mean %.2f p50 %d p90 %d p99 %d max %d" % (piv.mean(), *np.percentile(piv, [50, 90, 99]).astype(int), piv.max()))
big = o[piv >= 40]
print("steps with >= 40 pivots:", len(big))
for r in big[:25]:
    print("   n_blocks %2d itf %2d pivots %3d newton %2d status %02x stable %d su %d" % (
        r["n_blocks"], r["n_interfaces"], r["lp_pivots"], r["newton_iters"], r["solver_status"], r["stable"], r["stable_unfrozen"]))
nw = o["newton_iters"]
print("newton iters per step mean %.3f; steps with a Newton solve %d of %d" % (nw.mean(), int((nw > 0).sum()), len(o)))
for nbk in range(1, 16):
    s = o[o["n_blocks"] == nbk]
    if len(s):
        print("   n_blocks %2d n %6d pivots mean %.2f p99 %d max %d | newton>0 %d | by lp %.2f" % (
            nbk, len(s), s["lp_pivots"].mean(), int(np.percentile(s["lp_pivots"], 99)), s["lp_pivots"].max(), int((s["newton_iters"] > 0).sum()),
            (((s["solver_status"] & 16) != 0).mean() + ((s["solver_status"] & 32) != 0).mean()) / 2))
