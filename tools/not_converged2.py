import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests.test_gpu_properties import _build, TOTAL
env = _build(0, TOTAL)
o = env.read_out()
nc = (o["solver_status"] & 3) != 0
print("not converged:", int(nc.sum()), "of", TOTAL)
idx = np.nonzero(nc)[0][:20]
for e in idx:
    print(e, "status", o["solver_status"][e], "nb", o["n_blocks"][e], "itf", o["n_interfaces"][e], "res %.3e %.3e" % (o["residual"][e], o["residual_unfrozen"][e]),
          "iters", o["newton_iters"][e], "stable", o["stable"][e], o["stable_unfrozen"][e], "mu", [0.3, 0.8, 2.0][e % 3])
it = o["newton_iters"]
print("iters mean %.1f p99 %d max %d" % (it.mean(), np.percentile(it, 99), it.max()))
np.save("gpurun_out/nc_idx.npy", np.nonzero(nc)[0])
blocks, nb = env.get_state()
np.save("gpurun_out/nc_blocks.npy", blocks[idx]); np.save("gpurun_out/nc_nb.npy", nb[idx])
