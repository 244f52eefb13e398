"""Cycle counts of the LP verdict path per env step (BW_PROFILE build):
BRIDGES_B200_LIB=<pkg>/libbridges_b200_prof.so python tools/lp_profile.py [E] [bridge|tower2|tower4]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bridges_b200.envs.batched import BatchedAssemblyGym
from bench import task_def, bridge_def, X_GROUND
E = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
CASE = sys.argv[2] if len(sys.argv) > 2 else "bridge"
if CASE == "bridge":
    env = BatchedAssemblyGym(E, ["shapes/trapezoid.urdf", "shapes/hexagon.urdf"], max_steps=15); env.reset(bridge_def(5)); AMAX = 1024
elif CASE == "tower4":
    env = BatchedAssemblyGym(E, ["shapes/trapezoid.urdf"], max_steps=15); env.reset(task_def(4)); AMAX = 256
else:
    env = BatchedAssemblyGym(E, ["shapes/trapezoid.urdf"], max_steps=10); env.reset(task_def(2)); AMAX = 128
img = torch.zeros((E, 1, 64, 64), dtype=torch.float32, device='cuda')
outs, subs = [], []
for i in range(60):
    env.enumerate_actions(X_GROUND, (0.0,), amax=AMAX, with_bits=False)
    acts, _ = env.select_random(seed=i)
    env.step(acts, block_img=img)
    if i >= 15:
        outs.append(env.read_out().copy()); subs.append(img[:, 0, 0, :48].cpu().numpy().copy())
    env.reset_done()
o = np.concatenate(outs); s = np.concatenate(subs).astype(np.float64)
tot = o["reward"].astype(np.float64)
per_launch_max = [r["reward"].astype(np.float64).max() for r in outs]
print("case", CASE, "E", E, "| per launch: max-env cycles mean %.0f, mean-env cycles %.0f" % (np.mean(per_launch_max), tot.mean()))
names = ["copy-in", "setup", "run", "copy-out", "duals+pricing", "column+ratio", "update", "certificate"]
lp = s[:, 32:40]; piv = s[:, 40]
phase = {"load+place+faces": o["distance_to_targets"][:, 0], "interfaces+contacts+adj": o["distance_to_targets"][:, 1],
         "solve phase (LP + fallback) warp0": o["distance_to_targets"][:, 2], "bookkeeping": o["residual"], "raster+lin": o["residual_unfrozen"]}
for k, v in phase.items():
    print("  %-36s mean %8.0f p99 %8.0f max %8.0f" % (k, v.mean(), np.percentile(v, 99), v.max()))
print("LP path, cycles per env step (pivots per step mean %.2f):" % piv.mean())
for k, nm in enumerate(names):
    print("  %-16s mean %8.0f p99 %8.0f max %8.0f" % (nm, lp[:, k].mean(), np.percentile(lp[:, k], 99), lp[:, k].max()))
pv = np.maximum(piv, 1)
print("per pivot: duals+pricing %.0f  column+ratio %.0f  update %.0f" % (lp[:, 4].sum() / piv.sum(), lp[:, 5].sum() / piv.sum(), lp[:, 6].sum() / piv.sum()))
order = np.argsort(-tot)[:15]
print("slowest env steps: total | lp copy-in setup run copy-out | pivots n_blocks itf newton status")
for k in order:
    print("  %8.0f | %7.0f %7.0f %7.0f %7.0f | %3d %2d %2d %2d %02x" % (tot[k], lp[k, 0], lp[k, 1], lp[k, 2], lp[k, 3], piv[k], o["n_blocks"][k],
          o["n_interfaces"][k], o["newton_iters"][k], o["solver_status"][k]))
