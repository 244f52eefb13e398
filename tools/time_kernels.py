"""Event timings of the individual kernels on the bench workload."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bridges_b200.envs.batched import BatchedAssemblyGym
from bench import task_def, X_GROUND
E = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
env = BatchedAssemblyGym(E, ["shapes/trapezoid.urdf"], max_steps=10)
env.reset(task_def(2))
img = torch.zeros((E, 1, 64, 64), device="cuda"); binary = torch.zeros((E, 6), device="cuda")
def ev(): return torch.cuda.Event(enable_timing=True)
T = {"enumerate(bits)": [], "enumerate(no bits)": [], "select": [], "step": [], "reset_done": [], "observe(block f32)": []}
for i in range(60):
    a, b = ev(), ev(); a.record(); env.enumerate_actions(X_GROUND, (0.0,), amax=128, with_bits=True); b.record()
    c, d = ev(), ev(); c.record(); env.enumerate_actions(X_GROUND, (0.0,), amax=128, with_bits=False); d.record()
    e_, f = ev(), ev(); e_.record(); acts, _ = env.select_random(seed=i); f.record()
    g, h = ev(), ev(); g.record(); env.step(acts, block_img=img, binary=binary); h.record()
    k, l = ev(), ev(); k.record(); env.reset_done(); l.record()
    m, n = ev(), ev(); m.record(); env.lib.bw_observe(env.handle, img.data_ptr(), None, None, None); n.record()
    torch.cuda.synchronize()
    if i >= 20:
        for key, (x, y) in zip(T, ((a, b), (c, d), (e_, f), (g, h), (k, l), (m, n))):
            T[key].append(x.elapsed_time(y) * 1e3)
for k, v in T.items():
    print(f"{k:22s} mean {np.mean(v):8.1f} us  min {np.min(v):8.1f} us")
