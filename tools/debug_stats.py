"""Scratch diagnostics: distribution of step outputs on the bench workload."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bridges_b200.envs.batched import BatchedAssemblyGym
from bench import task_def, X_GROUND
E = 1024
env = BatchedAssemblyGym(E, ["shapes/trapezoid.urdf"], max_steps=10)
env.reset(task_def(2))
for i in range(30):
    env.enumerate_actions(X_GROUND, (0.0,), amax=128, with_bits=False)
    acts, idx = env.select_random(seed=i)
    env.step(acts)
    out = env.read_out()
    if i % 5 == 0:
        for f in ("stable", "stable_unfrozen", "terminated", "truncated", "error", "solver_status", "n_blocks"):
            print(i, f, np.unique(out[f], return_counts=True))
        print(i, "idx<0:", int((idx.cpu().numpy() < 0).sum()), "iters mean", out["newton_iters"].mean(), "max", out["newton_iters"].max())
    env.reset_done()
print("---- with fused observation outputs")
block_img = torch.empty((E, 1, 64, 64), dtype=torch.float32, device="cuda")
binary = torch.empty((E, 6), dtype=torch.float32, device="cuda")
for i in range(12):
    env.enumerate_actions(X_GROUND, (0.0,), amax=128, with_bits=False)
    acts, idx = env.select_random(seed=100 + i)
    env.step(acts, block_img=block_img, binary=binary)
    snap = env._out.clone()
    out = snap.cpu().numpy().view(env.dt["step_out"])
    if i % 4 == 0:
        for f in ("stable", "stable_unfrozen", "n_blocks"):
            print(i, f, np.unique(out[f], return_counts=True))
        bits, _ = env.raster_bits()
        img = block_img.cpu().numpy()[:, 0]
        print(i, "img == bits:", np.array_equal(img.astype(bool), env.bits_to_bool(bits)), "binary[:,0]==stable:",
              np.array_equal(binary.cpu().numpy()[:, 0], out["stable"].astype(np.float32)))
    env.reset_done()
