"""BASELINE.json configs[4] in small: horizontal-bridge task with a mixed trapezoid + hexagon library,
env-parallel over the ranks of one box, transitions all-gathered into every rank's replay memory over
NCCL.  Launch: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/multi_gpu_rollout.py"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from bridges_b200.envs.batched import BatchedAssemblyGym
from bridges_b200.rollout import DeviceReplayBuffer, random_policy, rollout_lockstep

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
E, n_obst, steps = 1024, 5, 40
env = BatchedAssemblyGym(E, ["shapes/trapezoid.urdf", "shapes/hexagon.urdf"], max_steps=15, device=local)
obstacles = [(i * 0.6, 0, 0.3) for i in range(1, n_obst + 1)]
env.reset(dict(obstacles=obstacles, targets=[(n_obst * 0.6 + 1.5, 0, 0.3)]))      # gym_env.py:36-40
replay = DeviceReplayBuffer(world * E * steps, dev)
xg = np.linspace(-2, 0, 10)
rollout_lockstep(env, random_policy(seed=100 + rank), 4, xg, amax=512, replay=None)       # warm-up
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
rollout_lockstep(env, random_policy(seed=rank), steps, xg, amax=512, replay=replay)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
data = {k: v[:len(replay)] for k, v in replay.data.items()}
envs_seen = int(torch.unique(data["env"]).numel())
ok = bool(torch.equal(data["next_block_bits"], data["block_bits"] | data["action_bits"]))
t = torch.tensor([dt], device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # every rank holds the same replay content
    digest = torch.stack([data["block_bits"].sum(), data["action_bits"].sum(), data["reward"].sum().long()]).long()
    all_d = [torch.zeros_like(digest) for _ in range(world)]
    dist.all_gather(all_d, digest)
    same = all(bool(torch.equal(all_d[0], d)) for d in all_d)
else:
    same = True
if rank == 0:
    print(json.dumps(dict(ranks=world, envs_per_rank=E, steps=steps, transitions_in_replay=len(replay),
                          envs_seen=envs_seen, rasters_consistent=ok, replay_identical_on_all_ranks=same,
                          rollout_env_steps_per_s=world * E * steps / float(t.item()),
                          done_frac=float(data["done"].float().mean()), mean_reward=float(data["reward"].mean()))))
if world > 1:
    dist.destroy_process_group()
