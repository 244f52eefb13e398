"""BASELINE.json configs[4] in small: horizontal-bridge task with a mixed trapezoid + hexagon library, env-parallel
over the ranks of one box, fused rollout (`bw_rollout_random`), every T-step chunk of packed records all-gathered
into every rank's replay ring over NCCL on a side stream.
Launch: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/multi_gpu_rollout.py [E] [task]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from bench import X_GROUND, bridge_def, task_def
from bridges_b200.envs.batched import BatchedAssemblyGym
from bridges_b200.rollout import REC, FusedRollout, TransitionRing

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
E = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
task = sys.argv[2] if len(sys.argv) > 2 else "bridge"
T, chunks = 16, 12
if task == "bridge":
    env = BatchedAssemblyGym(E, ["shapes/trapezoid.urdf", "shapes/hexagon.urdf"], max_steps=15, device=local)
    env.reset(bridge_def(5))
    amax = 1024
else:
    env = BatchedAssemblyGym(E, ["shapes/trapezoid.urdf"], max_steps=10, device=local)
    env.reset(task_def(2))
    amax = 128
ring = TransitionRing(4 * world * T * E, dev)
roll = FusedRollout(env, X_GROUND, (0.0,), amax=amax, chunk_steps=T, ring=ring)
roll.collect_random(2, seed=100)                    # warm-up
roll.drain()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
ev0.record()
roll.collect_random(chunks, seed=7)
roll.drain()
ev1.record()
torch.cuda.synchronize()
wall = time.perf_counter() - t0
dev_s = ev0.elapsed_time(ev1) * 1e-3
valid = ring.column("valid")
n_valid = int((valid != 0).sum())
rec = ring.numpy()
v = rec[rec["valid"] == 1]
ok = bool(np.array_equal(v["next_block_bits"], v["block_bits"] | v["action_bits"]))
t = torch.tensor([dev_s, wall], device=dev)
same = True
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    digest = torch.tensor([int(rec["env"].astype(np.int64).sum()), int(rec["step"].astype(np.int64).sum()),
                           int(rec["block_bits"][:, 32].astype(np.int64).sum() & 0x7fffffff)], device=dev)
    all_d = [torch.zeros_like(digest) for _ in range(world)]
    dist.all_gather(all_d, digest)
    same = all(bool(torch.equal(all_d[0], d)) for d in all_d)
if rank == 0:
    iters = chunks * T
    print(json.dumps(dict(ranks=world, envs_per_rank=E, task=task, iterations=iters, chunk_steps=T,
                          transitions_per_s=world * E * iters / float(t[0].item()),
                          transitions_per_s_per_gpu=E * iters / float(t[0].item()),
                          ms_per_iteration=1e3 * float(t[0].item()) / iters, wall_s=float(t[1].item()),
                          record_bytes=REC, gather_bytes_per_chunk_per_rank=(world - 1) * T * E * REC,
                          ring_records=int(ring.size), valid_frac=n_valid / max(int(ring.size), 1),
                          envs_seen=int(np.unique(rec["env"]).size), rasters_consistent=ok,
                          ring_identical_on_all_ranks=same, done_frac=float(v["done"].mean()),
                          mean_reward=float(v["reward"].mean()))))
if world > 1:
    dist.destroy_process_group()
