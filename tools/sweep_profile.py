"""Per-phase cycle counts of the step kernel on the 65,536-assembly sweep (needs `make -C csrc prof`).
run: BRIDGES_B200_LIB=<pkg>/libbridges_b200_prof.so python tools/sweep_profile.py [N]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bridges_b200.envs.batched import BatchedAssemblyGym
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
env = BatchedAssemblyGym(n, ["shapes/trapezoid.urdf", "shapes/hexagon.urdf", "shapes/cube1.urdf"], max_steps=None)
ids = np.arange(n)
env.set_mu(np.array([0.3, 0.8, 2.0])[ids % 3])
env.reset(dict())
rng = np.random.default_rng(0)
target = rng.integers(1, 16, size=n)
for k in range(15):
    env.enumerate_actions(np.linspace(-2.0, 4.0, 13), (0.0, 0.25, -0.25), amax=1024, with_bits=False)
    acts, _ = env.select_random(seed=12345 + k)
    env.step(acts, mask=(target > k).astype(np.uint8))
if len(sys.argv) > 2 and sys.argv[2] == "noprof":
    # plain build, for an ncu capture of one pass: reset's evaluation + 15 build steps + these = launch 18 is a warm pass
    for _ in range(3):
        env.evaluate()
    env.sync()
    print("3 evaluation passes done")
    sys.exit(0)
img = torch.zeros((n, 1, 64, 64), dtype=torch.float32, device="cuda")
env.evaluate(block_img=img)
o = env.read_out().copy()
sb = img[:, 0, 0, :16].cpu().numpy()
names = ["load+place+faces", "interfaces+contacts+adj", "solve warp0", "solve warp1", "bookkeeping", "raster", "total"]
cols = [o["distance_to_targets"][:, 0], o["distance_to_targets"][:, 1], o["distance_to_targets"][:, 2],
        o["distance_to_targets"][:, 3], o["residual"], o["residual_unfrozen"], o["reward"].astype(np.float64)]
for nm, c in zip(names, cols):
    print(f"{nm:28s} mean {c.mean():9.0f}  p50 {np.percentile(c,50):9.0f}  p99 {np.percentile(c,99):9.0f}  max {c.max():9.0f} cycles")
it = o["newton_iters"].astype(float)
print("newton iters mean %.2f max %d" % (it.mean(), it.max()), "status bits", np.unique(o["solver_status"], return_counts=True))
for nb in range(1, 16):
    sel = o["n_blocks"] == nb
    if sel.sum():
        print(f"n_blocks={nb:2d} n={sel.sum():5d} total mean {cols[6][sel].mean():9.0f} max {cols[6][sel].max():9.0f} itf mean {o['n_interfaces'][sel].mean():5.1f} iters mean {it[sel].mean():5.1f}"
              f" | w0 {cols[2][sel].mean():8.0f} w1 {cols[3][sel].mean():8.0f} itf-phase {cols[1][sel].mean():8.0f} screen(w1) {sb[sel, 5].mean():8.0f}")
for k, nm in enumerate(["grad", "assemble H", "cholesky+solves", "A^T d + dots", "line search + update", "mechanism screen"]):
    print(f"warp1 solve / {nm:24s} mean {sb[:, k].mean():9.0f}  share {sb[:, k].sum() / sb[:, :6].sum():6.1%}")
