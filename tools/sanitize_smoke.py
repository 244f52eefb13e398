"""Small end-to-end run for compute-sanitizer (memcheck): every kernel, small sizes, 16-block capacity."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bridges_b200.envs.batched import BatchedAssemblyGym
xg = np.linspace(-2, 0, 10)
for max_steps, shapes in ((10, ["shapes/trapezoid.urdf"]), (None, ["shapes/trapezoid.urdf", "shapes/hexagon.urdf", "shapes/cube1.urdf"])):
    env = BatchedAssemblyGym(48, shapes, max_steps=max_steps)
    env.reset(dict(obstacles=[(0.6, 0, 0.3)], targets=[(0.6, 0, 0.9)]))
    img = torch.zeros((48, 1, 64, 64), device="cuda"); u8 = torch.zeros((48, 64, 64), dtype=torch.uint8, device="cuda")
    binary = torch.zeros((48, 6), device="cuda")
    for i in range(18):
        env.enumerate_actions(xg, (0.0, 0.25), amax=512)
        acts, _ = env.select_random(seed=i)
        env.step(acts, block_img=img, binary=binary, block_u8=u8)
        out = env.read_out()
        if max_steps:
            env.reset_done()
    env.observe(block=True, binary=True, obstacle=True, reward=True)
    env.get_forces(0); env.get_forces(1); env.get_state(); env.raster_bits()
    print("ok", max_steps, "max blocks", int(out["n_blocks"].max()), "errors", np.unique(out["error"]))
    env.close()
