import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bridges_b200.envs.batched import BatchedAssemblyGym
from bench import task_def, X_GROUND
E = 1024
def run(with_obs, seeds):
    env = BatchedAssemblyGym(E, ["shapes/trapezoid.urdf"], max_steps=10)
    env.reset(task_def(2))
    block_img = torch.zeros((E, 1, 64, 64), dtype=torch.float32, device="cuda")
    binary = torch.full((E, 6), -7.0, dtype=torch.float32, device="cuda")
    for i, sd in enumerate(seeds):
        env.enumerate_actions(X_GROUND, (0.0,), amax=128, with_bits=False)
        acts, idx = env.select_random(seed=sd)
        if with_obs:
            env.step(acts, block_img=block_img, binary=binary)
        else:
            env.step(acts)
        out = env.read_out()
        if i % 6 == 5:
            print(with_obs, i, "stable", int(out["stable"].sum()), "unfrozen", int(out["stable_unfrozen"].sum()),
                  "mean blocks %.2f" % out["n_blocks"].mean(), "valid idx<0", int((idx.cpu().numpy() < 0).sum()))
            if with_obs:
                b = binary.cpu().numpy()
                print("   binary[:6]", b[:6, 0], "stable[:6]", out["stable"][:6], "mismatch", int((b[:, 0] != out["stable"]).sum()),
                      "other cols nonzero", int((b[:, 1:] != 0).sum()))
        env.reset_done()
    env.close()
run(False, range(100, 130))
run(True, range(100, 130))
run(False, range(0, 30))
