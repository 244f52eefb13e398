"""Experiment: the 1,024 environments of one GPU as k lock-step groups (k handles, k CUDA streams), so that one group's
candidate stage runs while another group's step kernel waits for its slowest environment.
usage: python tools/pipelined_rollout.py [envs] [groups ...]      (transitions/s of bw_rollout_random per setting)"""
import sys, time
import torch
sys.path.insert(0, ".")
from bridges_b200.envs.batched import BatchedAssemblyGym
from bridges_b200.rollout import FusedRollout
import bench

E_total = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
settings = [int(a) for a in sys.argv[2:]] or [1, 2, 4]
task = bench.bridge_def(5)                       # the headline task of bench.py
T, chunks = 16, 8
for k in settings:
    E = E_total // k
    streams = [torch.cuda.Stream() for _ in range(k)]
    rolls = []
    for g, st in enumerate(streams):
        with torch.cuda.stream(st):
            env = BatchedAssemblyGym(E, ["shapes/trapezoid.urdf", "shapes/hexagon.urdf"], max_steps=15, device=0)
            env.reset(dict(obstacles=task["obstacles"], targets=task["targets"]))
            amax = min(1024, max(128, 64 * ((env.max_candidates(len(bench.X_GROUND), 1) + 63) // 64)))
            rolls.append(FusedRollout(env, bench.X_GROUND, (0.0,), amax=amax, chunk_steps=T, ring=None))
    def run(n, seed):
        for c in range(n):
            for g, st in enumerate(streams):
                with torch.cuda.stream(st):
                    rolls[g].collect_random(1, seed=seed + 1000 * g + c)
    run(3, 100)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run(chunks, 7)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    iters = chunks * T
    print(f"groups {k} x {E} envs: {E * k * iters / dt / 1e6:.3f} M transitions/s, {1e3 * dt / iters:.4f} ms per iteration of all groups", flush=True)
    for r in rolls:
        r.env.close()
