"""Fused lock-step rollouts and a device-resident replay memory of packed records.

Batched counterpart of `rollout_episode` (robotoddler/training/successor_dqn.py:365-475) and of
`ReplayBuffer` / `PrioritizedReplayBuffer` (robotoddler/utils/replay_memory.py:10-93).  One rollout iteration is
a fixed sequence of kernels behind ONE C-ABI call (`bw_rollout_random` for the synthetic policy, `bw_rollout_begin`
/ `bw_rollout_commit` around a caller's policy): pick -> step (+ stabilities_freezing, lin_reward) -> record ->
auto-reset -> candidates of the next states -> "no candidate left" ends the episode (successor_dqn.py:409-411).
There is no torch operation on the per-step path; torch is used for device memory, for the learner-side sampling
and for the one collective per T-step chunk that copies freshly collected records into every rank's ring
(`all_gather_into_tensor` on a side stream, overlapped with the next chunk; NCCL on the GPU box, gloo in the CPU
tests).  A transition is a 1,608-byte `bw_transition` (three bit-packed rasters + scalars) instead of 3 x 16 KB
float images; `TransitionRing.sample` expands sampled records into the learner's float tensors with one kernel.
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import lib as L

IMG = 64
REC = 1608                               # sizeof(bw_transition)


def record_dtype():
    return L.np_dtypes()["transition"]


def _world(group=None):
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def _rank(group=None):
    return dist.get_rank(group) if dist.is_available() and dist.is_initialized() else 0


_TORCH_OF = {"<f4": torch.float32, "<f8": torch.float64, "<i4": torch.int32, "|u1": torch.uint8, "<u8": torch.int64}


def record_column(buf, name):
    """Field `name` of packed records buf (uint8 [..., REC]) as a tensor [..., *field shape] (one-byte fields are
    strided views, wider ones copies of their column; uint64 rasters come back as int64 bit patterns)."""
    fdt, off = record_dtype().fields[name][:2]
    if fdt.names:
        raise L.BridgesError("structured fields (action) are read through numpy: TransitionRing.numpy()")
    base = fdt.base if fdt.shape else fdt
    lead = buf.shape[:-1]
    col = buf[..., off:off + fdt.itemsize]
    if base.itemsize == 1:
        return col.reshape(lead + tuple(fdt.shape))
    return col.contiguous().view(_TORCH_OF[base.str]).reshape(lead + tuple(fdt.shape))


def gather_records(chunk, out, group=None, async_op=False):
    """chunk: uint8 [n, REC] of this rank; out: uint8 [world * n, REC], rank-major.  One collective."""
    if _world(group) == 1:
        out.copy_(chunk)
        return None
    return dist.all_gather_into_tensor(out, chunk, group=group, async_op=async_op)


class TransitionRing:
    """Ring of packed `bw_transition` records in device memory: `ReplayBuffer(capacity)` of the reference
    (a deque with maxlen).  Records are written in chunks (a rollout chunk, or the all-gathered chunks of all
    ranks); records whose `valid` byte is 0 (an environment that had no candidate in that iteration) stay in the
    ring but are never sampled.  With `prioritized=True` sampling follows `PrioritizedReplayBuffer`
    (replay_memory.py:45-93): probability proportional to |td_error| + 1e-5, new records enter with `td_error`."""

    def __init__(self, capacity, device, prioritized=False):
        self.capacity = int(capacity)
        self.device = torch.device(device)
        self.buf = torch.zeros((self.capacity, REC), dtype=torch.uint8, device=self.device)
        self.head = 0                    # next record to be written
        self.size = 0                    # records written so far, at most capacity
        self.dt = record_dtype()
        self.priorities = torch.zeros(self.capacity, dtype=torch.float32, device=self.device) if prioritized else None

    def __len__(self):
        """Number of sampleable transitions (host synchronisation)."""
        return int(self.valid_mask().sum().item())

    def reserve(self, n):
        """Region of n consecutive records for an in-place writer (rollout kernels, all-gather).  The ring
        advances; a region never wraps (capacity must be a multiple of the chunk sizes in use)."""
        if n > self.capacity or self.capacity % n != 0 or self.head % n != 0:
            raise L.BridgesError("ring capacity / head must be multiples of the chunk size")
        start = self.head
        self.head = (self.head + n) % self.capacity
        self.size = min(self.capacity, self.size + n)
        if self.priorities is not None:
            self.priorities[start:start + n] = 1e-5
        return start, self.buf[start:start + n]

    def push(self, records, td_error=None):
        """Append uint8 [n, REC] records (any n, wraps around)."""
        n = records.shape[0]
        if n == 0:
            return
        if n > self.capacity:
            records, n = records[-self.capacity:], self.capacity
            td_error = td_error[-self.capacity:] if td_error is not None else None
        pos = (torch.arange(n, device=self.device) + self.head) % self.capacity
        self.buf.index_copy_(0, pos, records.to(self.device))
        if self.priorities is not None:
            pr = torch.full((n,), 1e-5, device=self.device) if td_error is None else td_error.to(self.device).abs().float() + 1e-5
            self.priorities.index_copy_(0, pos, pr)
        self.head = (self.head + n) % self.capacity
        self.size = min(self.capacity, self.size + n)

    def column(self, name):
        """Field `name` of every record [capacity, ...] (see `record_column`)."""
        return record_column(self.buf, name)

    def valid_mask(self):
        m = self.column("valid") != 0
        if self.size < self.capacity:
            m = m & (torch.arange(self.capacity, device=self.device) < self.size)
        return m

    def sample_indices(self, batch_size, generator=None):
        valid = self.valid_mask()
        if self.priorities is not None:
            p = torch.where(valid, self.priorities, torch.zeros_like(self.priorities))
            return torch.multinomial(p, batch_size, replacement=True, generator=generator)
        idx = valid.nonzero(as_tuple=False).flatten()
        if idx.numel() == 0:
            raise L.BridgesError("the ring holds no transition yet")
        pick = torch.randint(0, idx.numel(), (batch_size,), device=self.device, generator=generator)
        return idx[pick]

    def update_priorities(self, indices, td_error):
        if self.priorities is not None:
            self.priorities[indices] = td_error.abs().float().to(self.device) + 1e-5

    def sample(self, env, batch_size, generator=None, indices=None):
        """`ReplayBuffer.sample(batch_size, stack_tensors=True)`: the sampled records expanded by
        `bw_unpack_transitions` into block / action / next_block images [B,1,64,64] f32, binary / next_binary
        [B,6], reward, lin_reward [B] f32, done [B] bool (plus `indices`)."""
        idx = self.sample_indices(batch_size, generator) if indices is None else indices
        idx = idx.to(torch.int64).contiguous()
        B, dev = idx.numel(), self.device
        img = lambda: torch.empty((B, 1, IMG, IMG), dtype=torch.float32, device=dev)
        out = dict(block_features=img(), action_features=img(), next_block_features=img(),
                   binary_features=torch.empty((B, 6), dtype=torch.float32, device=dev),
                   next_binary_features=torch.empty((B, 6), dtype=torch.float32, device=dev),
                   reward=torch.empty(B, dtype=torch.float32, device=dev),
                   lin_reward=torch.empty(B, dtype=torch.float32, device=dev),
                   done=torch.empty(B, dtype=torch.uint8, device=dev))
        env._check(env.lib.bw_unpack_transitions(
            env.handle, self.buf.data_ptr(), idx.data_ptr(), B, out["block_features"].data_ptr(),
            out["action_features"].data_ptr(), out["next_block_features"].data_ptr(), out["binary_features"].data_ptr(),
            out["next_binary_features"].data_ptr(), out["reward"].data_ptr(), out["lin_reward"].data_ptr(),
            out["done"].data_ptr()))
        out["done"] = out["done"].bool()
        out["indices"] = idx
        return out

    def numpy(self):
        """All written records as a numpy structured array (host copy; tests and checkpoints)."""
        n = self.size
        return self.buf[:n].cpu().numpy().reshape(-1).view(self.dt)[:n]


class _DevArray:
    """Zero-copy torch view of handle-owned device memory (`__cuda_array_interface__`)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = dict(shape=tuple(shape), typestr=typestr, data=(int(ptr), False), version=2)


class FusedRollout:
    """Drives `bw_rollout_*` for one `BatchedAssemblyGym` (one rank).

    collect_random(n_chunks): chunks of `chunk_steps` iterations with the built-in uniformly random policy, one
    C call per chunk; every finished chunk is all-gathered into the ring of every rank on a side stream while the
    next chunk runs.  step(index) / candidates(): one iteration around a caller's policy (a Q-network)."""

    def __init__(self, env, x_discr_ground, offset_values=(0.0,), amax=128, chunk_steps=16, ring=None, group=None):
        self.env, self.group = env, group
        self.E, self.T = env.num_envs, int(chunk_steps)
        self.world, self.rank = _world(group), _rank(group)
        g = np.ascontiguousarray(np.asarray(x_discr_ground, dtype=np.float64))
        o = np.ascontiguousarray(np.asarray(offset_values, dtype=np.float64))
        env._check(env.lib.bw_rollout_configure(env.handle, g.ctypes.data, g.size, o.ctypes.data, o.size, int(amax),
                                                self.rank * self.E))
        self.amax = int(amax)
        self.ring = ring
        n = self.T * self.E
        # two chunk buffers: the rollout fills one while the other is being gathered
        self.chunks = [torch.zeros((n, REC), dtype=torch.uint8, device=env.device) for _ in range(2)]
        self.slots = torch.zeros((self.E, REC), dtype=torch.uint8, device=env.device)      # step() records
        self.side = torch.cuda.Stream(device=env.device) if env.device.type == "cuda" else None
        self.pending = [None, None]      # event after which chunk buffer k may be overwritten
        self.n_chunk = 0
        self.gathered_bytes = 0
        self._view = None
        self.begin()

    # ---------------------------------------------------------------- candidates / caller's policy
    def begin(self):
        view = L.bw_rollout_view()
        self.env._check(self.env.lib.bw_rollout_begin(self.env.handle, C.byref(view)))
        E, A = self.E, self.amax
        dev = self.env.device
        wrap = lambda ptr, shape, ts: torch.as_tensor(_DevArray(ptr, shape, ts), device=dev)
        env = self.env
        if view.action_bits:             # no candidate store: dense raster copies
            bits, slot = wrap(view.action_bits, (E, A, IMG), "<i8"), None
        else:                            # the rasters stay in the handle's store: gathered on demand
            from .envs.batched import CandidateRasters
            slot = wrap(view.slot, (E, A), "<i4")
            bits = CandidateRasters(env, E, A, lambda e_ptr, i_ptr, n, out: env.lib.bw_rollout_gather_bits(
                env.handle, e_ptr, i_ptr, n, out))
        self._view = dict(amax=A, cand=wrap(view.cand, (E * A * 40,), "|u1"), valid=wrap(view.valid, (E, A), "|u1"),
                          n=wrap(view.n_cand, (E,), "<i4"), n_valid=wrap(view.n_valid, (E,), "<i4"), bits=bits, slot=slot)
        return self._view

    def candidates(self):
        """Candidate buffers of the current states (torch views of the handle's memory, refreshed in place by
        every iteration): cand bytes [E*amax*40], valid u8 [E,amax], n / n_valid i32 [E]; bits[env_idx, cand_idx] ->
        i64 [n,64] rasters of chosen candidates (a `CandidateRasters` over the handle's candidate store, or the dense
        i64 [E,amax,64] tensor when the store is switched off)."""
        return self._view

    def step(self, index, obs=None):
        """One iteration with the caller's choice: index int32 CUDA tensor [E].  Returns the records of this
        iteration (uint8 [E, REC], overwritten by the next call)."""
        index = index.to(torch.int32).contiguous()
        self.env._check(self.env.lib.bw_rollout_commit(self.env.handle, index.data_ptr(), self.slots.data_ptr(),
                                                       C.byref(obs) if obs is not None else None))
        self._keep = index
        return self.slots

    # ---------------------------------------------------------------- built-in random policy, chunked
    def collect_random(self, n_chunks, seed=0, gather=True):
        """n_chunks x chunk_steps iterations of every environment.  Returns (iterations, records written to the
        ring per rank and chunk)."""
        env, n = self.env, self.T * self.E
        main = torch.cuda.current_stream(env.device)
        for _ in range(n_chunks):
            k = self.n_chunk & 1
            if self.pending[k] is not None:
                main.wait_event(self.pending[k])              # the gather of this buffer's previous content is done
            buf = self.chunks[k]
            env._check(env.lib.bw_rollout_random(env.handle, self.T, int(seed) & (2 ** 64 - 1), buf.data_ptr(), n, 0))
            if self.ring is not None:
                self._publish(buf, k, gather, main)
            self.n_chunk += 1
        return n_chunks * self.T

    def _publish(self, buf, k, gather, main):
        """Copy / all-gather a finished chunk into the ring on the side stream."""
        world = self.world if gather else 1
        start, region = self.ring.reserve(world * buf.shape[0])
        ready = torch.cuda.Event()
        ready.record(main)
        with torch.cuda.stream(self.side):
            self.side.wait_event(ready)
            if world > 1:
                dist.all_gather_into_tensor(region, buf, group=self.group)
                self.gathered_bytes += (world - 1) * buf.numel()
            else:
                region.copy_(buf, non_blocking=True)
            done = torch.cuda.Event()
            done.record(self.side)
        self.pending[k] = done

    def drain(self):
        """Wait (stream-wise) until every published chunk is in the ring."""
        main = torch.cuda.current_stream(self.env.device)
        for ev in self.pending:
            if ev is not None:
                main.wait_event(ev)


def q_network_policy(policy_net, reward_features, obstacle_features, epsilon=0.0, seed=0, chunk_rows=8192):
    """Greedy / epsilon-greedy choice over the valid candidates of EVERY environment with one batched pass through
    an existing Q-network -- the lock-step form of the inference in `rollout_episode`
    (robotoddler/training/successor_dqn.py:383-390).  `policy_net` keeps the reference's signature
        policy_net(block_features, binary_features, action_features, reward_features, obstacle_features)
            -> (q_values [R], ...)
    (models/cv.py:76-105 SuccessorMLP, :41-65 ConvNet, ...) and sees R = sum of valid candidates rows: row r
    pairs the state of its environment with one candidate raster.  reward_features / obstacle_features:
    [E,1,64,64] task features (`env.observe(reward=True, obstacle=True)`).  Returns policy(env, cand) -> int32
    index tensor [E] (-1 where an environment has no valid candidate), the argument of `FusedRollout.step`."""
    gen = {"g": None}

    def policy(env, cand):
        E, dev, amax = env.num_envs, env.device, cand["amax"]
        if gen["g"] is None:
            gen["g"] = torch.Generator(device=dev)
            gen["g"].manual_seed(seed)
        valid = cand["valid"].bool() & (torch.arange(amax, device=dev)[None, :] < cand["n"][:, None])
        e_idx, a_idx = valid.nonzero(as_tuple=True)
        state = env.observe(block=True, binary=True)
        q_full = torch.full((E, amax), float("-inf"), device=dev)
        was_training = getattr(policy_net, "training", False)
        if hasattr(policy_net, "eval"):
            policy_net.eval()
        with torch.no_grad():
            for lo in range(0, e_idx.numel(), chunk_rows):
                er, ar = e_idx[lo:lo + chunk_rows], a_idx[lo:lo + chunk_rows]
                action_f = env.expand_bits(cand["bits"][er, ar].contiguous())      # only the valid candidates' rasters
                q = policy_net(state["block"][er], state["binary"][er], action_f, reward_features[er], obstacle_features[er])
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
        return torch.where(has, index, torch.full_like(index, -1)).to(torch.int32)
    return policy


def rollout_policy(roll, policy, n_steps, ring=None):
    """n_steps iterations of `roll` (a FusedRollout) around a caller's policy(env, cand) -> index; the records of
    every iteration are pushed to `ring` (local, not gathered).  Returns the records of the last iteration."""
    last = None
    for _ in range(n_steps):
        index = policy(roll.env, roll.candidates())
        last = roll.step(index)
        if ring is not None:
            ring.push(last)
    return last
