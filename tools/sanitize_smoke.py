"""Small end-to-end run for compute-sanitizer (memcheck): every kernel, small sizes, 16-block capacity."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bridges_b200.envs.batched import BatchedAssemblyGym
xg = np.linspace(-2, 0, 10)
for max_steps, shapes in ((10, ["shapes/trapezoid.urdf"]), (None, ["shapes/trapezoid.urdf", "shapes/hexagon.urdf", "shapes/cube1.urdf"])):
    env = BatchedAssemblyGym(48, shapes, max_steps=max_steps, collision=(max_steps is None))
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
    # host entry point: staged (pageable numpy buffers) and zero-copy (pinned buffers)
    import ctypes as C
    from bridges_b200 import lib as L
    env.enumerate_actions(xg, (0.0, 0.25), amax=512)
    acts, _ = env.select_random(seed=99)
    dt = env.dt
    pinned = dict(act=acts.cpu().pin_memory(), out=torch.zeros(48 * dt["step_out"].itemsize, dtype=torch.uint8).pin_memory(),
                  u8=torch.zeros((48, 64, 64), dtype=torch.uint8).pin_memory(), b=torch.zeros((48, 6)).pin_memory())
    obs = L.bw_obs_out(None, pinned["u8"].data_ptr(), pinned["b"].data_ptr())
    L.check(env.lib, env.handle, env.lib.bw_step_host(env.handle, pinned["act"].data_ptr(), None, pinned["out"].data_ptr(), C.byref(obs)))
    h_act = acts.cpu().numpy().copy(); h_out = np.zeros(48, dtype=dt["step_out"]); h_u8 = np.zeros((48, 64, 64), dtype=np.uint8)
    obs = L.bw_obs_out(None, h_u8.ctypes.data, None)
    L.check(env.lib, env.handle, env.lib.bw_step_host(env.handle, h_act.ctypes.data, None, h_out.ctypes.data, C.byref(obs)))
    env.observe(block=True, binary=True, obstacle=True, reward=True)
    env.get_forces(0); env.get_forces(1); env.get_state(); env.raster_bits()
    print("ok", max_steps, "max blocks", int(out["n_blocks"].max()), "errors", np.unique(out["error"]))
    env.close()
