"""Find assemblies of the sweep whose solve ended 'not converged' and print their residual history inputs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bridges_b200.envs.batched import BatchedAssemblyGym
n = 65536
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
env.evaluate()
out = env.read_out()
bad = np.nonzero(out["solver_status"])[0]
print("not converged:", bad, out["solver_status"][bad])
blocks, nb = env.get_state()
res = out["residual"]; resu = out["residual_unfrozen"]
print("residuals", res[bad], resu[bad], "iters", out["newton_iters"][bad], "n_blocks", nb[bad], "mu", np.array([0.3, 0.8, 2.0])[bad % 3])
band = ((res > 1e-9) & (res < 1e-4)) | ((resu > 1e-9) & (resu < 1e-4))
print("assemblies with a residual inside the (1e-9, 1e-4) band:", int(band.sum()), "of", n)
np.save("gpurun_out/not_converged_blocks.npy", blocks[bad])
for e in bad:
    print(e, [(float(b["x"]), float(b["z"]), float(b["c"]), float(b["s"]), int(b["shape"])) for b in blocks[e][:nb[e]]])
