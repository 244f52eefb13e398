"""Per-phase cycle counts of the step kernel on the bench workload (needs `make -C csrc prof`).
run: BRIDGES_B200_LIB=<pkg>/libbridges_b200_prof.so python tools/phase_profile.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bridges_b200.envs.batched import BatchedAssemblyGym
from bench import task_def, X_GROUND
E = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
env = BatchedAssemblyGym(E, ["shapes/trapezoid.urdf"], max_steps=10)
env.reset(task_def(2))
rows = []
subs = []
img = torch.zeros((E, 1, 64, 64), dtype=torch.float32, device='cuda')
for i in range(40):
    env.enumerate_actions(X_GROUND, (0.0,), amax=128, with_bits=False)
    acts, _ = env.select_random(seed=i)
    env.step(acts, block_img=img)
    out = env.read_out()
    if i >= 15:
        rows.append(out.copy())
        subs.append(img[:, 0, 0, :16].cpu().numpy().copy())
    env.reset_done()
o = np.concatenate(rows)
names = ["load+place+faces", "interfaces+contacts+adj", "solve warp0", "solve warp1", "bookkeeping", "raster", "total"]
cols = [o["distance_to_targets"][:, 0], o["distance_to_targets"][:, 1], o["distance_to_targets"][:, 2],
        o["distance_to_targets"][:, 3], o["residual"], o["residual_unfrozen"], o["reward"].astype(np.float64)]
for n, c in zip(names, cols):
    print(f"{n:28s} mean {c.mean():9.0f}  p50 {np.percentile(c,50):9.0f}  p99 {np.percentile(c,99):9.0f}  max {c.max():9.0f} cycles")
it = o["newton_iters"].astype(float)
solve = np.maximum(cols[2], cols[3])
print("newton iters mean %.1f max %d; cycles per newton iter (sum of both warps' time / iters): %.0f" %
      (it.mean(), it.max(), (cols[2] + cols[3]).sum() / max(it.sum(), 1)))
for nb in range(1, 11):
    sel = o["n_blocks"] == nb
    if sel.sum():
        print(f"n_blocks={nb:2d} n={sel.sum():5d} total mean {cols[6][sel].mean():9.0f} max {cols[6][sel].max():9.0f}  iters mean {it[sel].mean():5.1f}")

sb = np.concatenate(subs)
for k, n in enumerate(["grad(A^T y, proj, A f)", "assemble H", "cholesky+solves", "A^T d + dots", "line search + update", "mechanism screen"]):
    print(f"warp1 solve / {n:24s} mean {sb[:, k].mean():9.0f}  share {sb[:, k].sum() / sb[:, :6].sum():6.1%}")
for k, n in enumerate(["memset+task load", "targets+state write", "distances"]):
    print(f"bookkeeping cumulative / {n:22s} mean {sb[:, 8 + k].mean():9.0f}")
big = o["n_blocks"] >= 8
if big.sum():
    print(f"--- envs with >= 8 blocks (n={int(big.sum())}): warp1 solve cycles by sub-phase, mean per env")
    for k, n in enumerate(["grad", "assemble H", "cholesky+solves", "A^T d + dots", "line search + update", "mechanism screen"]):
        print(f"   {n:24s} {sb[big, k].mean():9.0f}  share {sb[big, k].sum() / sb[big, :6].sum():6.1%}")
    print("   total per env mean %.0f max %.0f; newton iters (both warps) mean %.1f" % (cols[6][big].mean(), cols[6][big].max(), it[big].mean()))
