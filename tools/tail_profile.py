"""Which environments set the length of the (single-wave) step launch?  Needs the BW_PROFILE build:
BRIDGES_B200_LIB=<pkg>/libbridges_b200_prof.so python tools/tail_profile.py [E]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bridges_b200.envs.batched import BatchedAssemblyGym
from bench import task_def, X_GROUND
from bench import bridge_def
E = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
CASE = sys.argv[2] if len(sys.argv) > 2 else "tower2"
if CASE == "bridge":       # BASELINE.json configs[4] shape: horizontal_bridge_setup(5), trapezoid + hexagon, max_steps 15
    env = BatchedAssemblyGym(E, ["shapes/trapezoid.urdf", "shapes/hexagon.urdf"], max_steps=15)
    env.reset(bridge_def(5))
    AMAX = 1024
elif CASE == "tower4":     # configs[2] shape
    env = BatchedAssemblyGym(E, ["shapes/trapezoid.urdf"], max_steps=15)
    env.reset(task_def(4))
    AMAX = 256
else:
    env = BatchedAssemblyGym(E, ["shapes/trapezoid.urdf"], max_steps=10)
    env.reset(task_def(2))
    AMAX = 128
print("case", CASE, "E", E)
rows = []
subs = []
img = torch.zeros((E, 1, 64, 64), dtype=torch.float32, device='cuda')
for i in range(60):
    env.enumerate_actions(X_GROUND, (0.0,), amax=AMAX, with_bits=False)
    acts, _ = env.select_random(seed=i)
    env.step(acts, block_img=img)
    out = env.read_out()
    if i >= 15:
        rows.append(out.copy())
        subs.append(img[:, 0, 0, :32].cpu().numpy().copy())
    env.reset_done()
per_launch_max = [r["reward"].astype(np.float64).max() for r in rows]
per_launch_mean = [r["reward"].astype(np.float64).mean() for r in rows]
print("per launch: max-env cycles mean %.0f  (mean-env cycles %.0f)" % (np.mean(per_launch_max), np.mean(per_launch_mean)))
o = np.concatenate(rows)
tot = o["reward"].astype(np.float64)
w0, w1 = o["distance_to_targets"][:, 2], o["distance_to_targets"][:, 3]
order = np.argsort(-tot)[:40]
print("slowest 40 env-steps: total w0 w1 | n_blocks n_itf iters | stable stable_unfrozen status")
for k in order:
    print("%8.0f %8.0f %8.0f | %2d %2d %3d | %d %d %d" % (tot[k], w0[k], w1[k], o["n_blocks"][k], o["n_interfaces"][k],
          o["newton_iters"][k], o["stable"][k], o["stable_unfrozen"][k], o["solver_status"][k]))
# the launch-setting env of each launch
print("launch-setting envs:")
for r in rows[:20]:
    t = r["reward"].astype(np.float64); k = int(np.argmax(t))
    print("  %8.0f w0 %8.0f w1 %8.0f | nb %2d itf %2d it %3d | s %d su %d" % (t[k], r["distance_to_targets"][k, 2], r["distance_to_targets"][k, 3],
          r["n_blocks"][k], r["n_interfaces"][k], r["newton_iters"][k], r["stable"][k], r["stable_unfrozen"][k]))
for name, sel in (("stable&stable_u", (o["stable"] == 1) & (o["stable_unfrozen"] == 1)), ("stable&!stable_u", (o["stable"] == 1) & (o["stable_unfrozen"] == 0)),
                  ("!stable", o["stable"] == 0)):
    if sel.sum():
        print("%-18s n=%6d  w0 mean %8.0f p99 %8.0f | w1 mean %8.0f p99 %8.0f | iters mean %.1f" % (name, sel.sum(), w0[sel].mean(), np.percentile(w0[sel], 99),
              w1[sel].mean(), np.percentile(w1[sel], 99), o["newton_iters"][sel].mean()))

sb = np.concatenate(subs)
print("warp 0 sub-phases of the 12 slowest env-steps, cycles per Newton step (grad, assemble, factor+solve, A^T d + dots, line search, | screen total):")
for k in order[:12]:
    it0 = max(sb[k, 22], 1.0)
    print("  nb %2d itf %2d it0 %3d it1 %3d | %s | screen %7.0f" % (o["n_blocks"][k], o["n_interfaces"][k], sb[k, 22], sb[k, 23],
          " ".join("%7.0f" % (sb[k, 16 + q] / it0) for q in range(5)), sb[k, 21]) +
          " | factor: dots %6.0f pivots %6.0f backsub %6.0f" % tuple(sb[k, 24 + q] / it0 for q in range(3)))
