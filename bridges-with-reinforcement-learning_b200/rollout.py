"""Lock-step rollouts and a device-resident replay memory.

Batched counterpart of `rollout_episode` (robotoddler/training/successor_dqn.py:365-475) and
`ReplayBuffer` (robotoddler/utils/replay_memory.py:10-43): every environment of a
`BatchedAssemblyGym` advances once per iteration, finished episodes are reset in place, and the
transitions stay on the GPU with their rasters bit-packed (512 B instead of 16 KB per image).
Across GPUs the only communication is the gather of freshly collected transitions into every
rank's replay memory (`torch.distributed.all_gather_into_tensor`, NCCL over NVLink on the GPU box,
gloo in the CPU tests); the environment step itself never communicates.
"""
import torch
import torch.distributed as dist

IMG = 64
# one transition = these fields (successor_dqn.py:27-44, minus the per-candidate next-state tensors,
# which are re-derived from next_block_bits by the candidate kernel when a batch is sampled)
FIELDS = (("block_bits", torch.int64, (IMG,)), ("action_bits", torch.int64, (IMG,)),
          ("next_block_bits", torch.int64, (IMG,)), ("binary", torch.float32, (6,)),
          ("next_binary", torch.float32, (6,)), ("reward", torch.float32, ()), ("lin_reward", torch.float32, ()),
          ("done", torch.bool, ()), ("env", torch.int32, ()))


def empty_batch(n, device):
    return {name: torch.zeros((n,) + shape, dtype=dtype, device=device) for name, dtype, shape in FIELDS}


def gather_transitions(batch, group=None):
    """Concatenate the per-rank batches of equal length on every rank (collective)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return batch
    world = dist.get_world_size(group)
    # one collective: every row is packed into a byte record, the records are gathered, then unpacked
    n = next(iter(batch.values())).shape[0]
    cols, layout = [], []
    for name, t in batch.items():
        src = (t.to(torch.uint8) if t.dtype == torch.bool else t).contiguous().reshape(n, -1)
        raw = src.view(torch.uint8)
        cols.append(raw)
        layout.append((name, t.dtype, tuple(t.shape[1:]), raw.shape[1]))
    packed = torch.cat(cols, dim=1).contiguous()
    dst = torch.empty((world * n, packed.shape[1]), dtype=torch.uint8, device=packed.device)
    dist.all_gather_into_tensor(dst, packed, group=group)
    out, off = {}, 0
    for name, dtype, shape, width in layout:
        raw = dst[:, off:off + width].contiguous()
        off += width
        if dtype == torch.bool:
            out[name] = raw.reshape((world * n,) + shape).to(torch.bool)
        else:
            out[name] = raw.view(dtype).reshape((world * n,) + shape)
    return out


class DeviceReplayBuffer:
    """Ring buffer of transitions in device memory (`push` / `sample`, replay_memory.py:10-43)."""

    def __init__(self, capacity, device):
        self.capacity = int(capacity)
        self.device = torch.device(device)
        self.data = empty_batch(self.capacity, self.device)
        self.size = 0
        self.head = 0

    def __len__(self):
        return self.size

    def push(self, batch, valid=None):
        """Append the rows of `batch` (optionally only those with `valid`)."""
        if valid is not None:
            idx = torch.nonzero(valid, as_tuple=False).flatten()
            batch = {k: v.index_select(0, idx) for k, v in batch.items()}
        n = next(iter(batch.values())).shape[0]
        if n == 0:
            return
        if n > self.capacity:
            batch = {k: v[-self.capacity:] for k, v in batch.items()}
            n = self.capacity
        pos = (torch.arange(n, device=self.device) + self.head) % self.capacity
        for name, t in batch.items():
            self.data[name].index_copy_(0, pos, t.to(self.device))
        self.head = (self.head + n) % self.capacity
        self.size = min(self.capacity, self.size + n)

    def sample(self, batch_size, generator=None):
        idx = torch.randint(0, self.size, (batch_size,), device=self.device, generator=generator)
        return {name: t.index_select(0, idx) for name, t in self.data.items()}


def random_policy(seed=0):
    """Uniformly random valid candidate (the synthetic policy of the benchmarks)."""
    state = {"step": 0}

    def policy(env, cand):
        actions, index = env.select_random(seed * 1000003 + state["step"] * 7919, cand)
        state["step"] += 1
        return actions, index
    return policy


def q_network_policy(policy_net, reward_features, obstacle_features, epsilon=0.0, seed=0, chunk_rows=8192):
    """Greedy / epsilon-greedy policy over the valid candidates of EVERY environment with one batched
    pass through an existing Q-network -- the lock-step form of the inference in `rollout_episode`
    (robotoddler/training/successor_dqn.py:383-390).  `policy_net` keeps the reference's signature
        policy_net(block_features, binary_features, action_features, reward_features, obstacle_features)
            -> (q_values [R], ...)
    (models/cv.py:76-105 SuccessorMLP, :41-65 ConvNet, ...) and sees R = sum of valid candidates rows:
    row r pairs the state of its environment with one candidate raster.  reward_features /
    obstacle_features: [E,1,64,64] task features (`env.observe(reward=True, obstacle=True)`).
    Exploration picks a uniformly random valid candidate with probability epsilon per environment."""
    gen = {"g": None}

    def policy(env, cand):
        E, dev, amax = env.num_envs, env.device, cand["amax"]
        if gen["g"] is None:
            gen["g"] = torch.Generator(device=dev)
            gen["g"].manual_seed(seed)
        valid = cand["valid"].bool()
        slots = torch.arange(amax, device=dev)[None, :] < cand["n"][:, None]
        valid = valid & slots
        e_idx, a_idx = valid.nonzero(as_tuple=True)
        state = env.observe(block=True, binary=True)
        q_full = torch.full((E, amax), float("-inf"), device=dev)
        was_training = getattr(policy_net, "training", False)
        if hasattr(policy_net, "eval"):
            policy_net.eval()
        with torch.no_grad():
            for lo in range(0, e_idx.numel(), chunk_rows):
                er, ar = e_idx[lo:lo + chunk_rows], a_idx[lo:lo + chunk_rows]
                action_f = env.expand_bits(cand["bits"][er, ar].contiguous())
                q = policy_net(state["block"][er], state["binary"][er], action_f, reward_features[er],
                               obstacle_features[er])
                q = q[0] if isinstance(q, (tuple, list)) else q
                q_full[er, ar] = q.reshape(-1).float()
        if was_training:
            policy_net.train()
        index = q_full.argmax(dim=1)
        if epsilon > 0.0:
            noise = torch.rand((E, amax), device=dev, generator=gen["g"]).masked_fill(~valid, -1.0)
            explore = torch.rand(E, device=dev, generator=gen["g"]) < epsilon
            index = torch.where(explore, noise.argmax(dim=1), index)
        has = valid.any(dim=1)
        item = env.dt["action"].itemsize
        chosen = cand["cand"].view(E, amax, item)[torch.arange(E, device=dev), index]
        noop = torch.from_numpy(env.actions_array([None])[:1].view("uint8").copy()).to(dev)
        actions = torch.where(has[:, None], chosen, noop[None, :].expand(E, item)).contiguous().view(-1)
        return actions, torch.where(has, index, torch.full_like(index, -1)).to(torch.int32)
    return policy


def rollout_lockstep(env, policy, n_steps, x_discr_ground, offset_values=(0.0,), amax=128, replay=None,
                     gather=True):
    """Advance every environment `n_steps` times.

    policy(env, cand) -> (uint8 CUDA tensor holding bw_action[E], int32 index tensor [E] into the
    candidates, -1 = no valid candidate).  Returns the last gathered batch; with `replay` the
    (gathered) transitions are pushed as they are produced.
    """
    E, dev = env.num_envs, env.device
    binary = torch.zeros((E, 6), dtype=torch.float32, device=dev)
    binary[:, 0] = 1.0                                          # empty scene: stable
    rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    env_ids = torch.arange(E, dtype=torch.int32, device=dev) + rank * E
    last = None
    fresh = torch.tensor([1.0, 0, 0, 0, 0, 0], device=dev).expand(E, 6)
    for _ in range(n_steps):
        cand = env.enumerate_actions(x_discr_ground, offset_values, amax=amax, with_bits=True)
        actions, index = policy(env, cand)
        has_action = index >= 0
        before = env.raster_bits_device()
        sel = cand["bits"][torch.arange(E, device=dev), index.clamp(min=0).long()]
        sel = torch.where(has_action[:, None], sel, torch.zeros_like(sel))
        next_binary = torch.empty_like(binary)
        out_dev = env.step(actions, binary=next_binary)
        after = env.raster_bits_device()
        out = env.out_fields(out_dev, ("reward", "lin_reward", "terminated", "truncated"))   # device views, no sync
        done = ((out["terminated"] | out["truncated"]) != 0) | ~has_action
        batch = dict(block_bits=before, action_bits=sel, next_block_bits=after, binary=binary.clone(),
                     next_binary=next_binary, reward=out["reward"].clone(), lin_reward=out["lin_reward"].clone(),
                     done=done, env=env_ids)
        keep = has_action                                   # envs without a valid candidate yield no transition
        if gather:
            full = gather_transitions(dict(batch, keep=keep))
            keep = full.pop("keep")
            batch = full
        if replay is not None:
            replay.push(batch, valid=keep)
        last = (batch, keep)
        env.reset_done()
        # environments that could not move end their episode too (rollout_episode, successor_dqn.py:409-411)
        env.reset(None, mask=(~has_action).to(torch.uint8))
        binary = torch.where(done[:, None], fresh, next_binary)
    return last
