"""Replay memory and transition gather: gloo world_size 2 on CPU for
the collective, GPU test for the lock-step rollout itself."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    from bridges_b200.rollout import DeviceReplayBuffer, empty_batch, gather_transitions
    dist.init_process_group("gloo", rank=rank, world_size=world)
    batch = empty_batch(4, "cpu")
    batch["reward"] += rank + 1
    batch["env"] += torch.arange(4, dtype=torch.int32) + 4 * rank
    batch["done"][rank] = True
    batch["block_bits"][:, 0] = rank + 10
    full = gather_transitions(batch)
    replay = DeviceReplayBuffer(6, "cpu")
    replay.push(full, valid=full["env"] % 2 == 0)            # 4 of the 8 gathered rows
    replay.push(full, valid=full["env"] >= 5)                # 3 more: wraps around the capacity of 6
    torch.save(dict(full=full, size=len(replay), head=replay.head, env=replay.data["env"].clone()),
               os.path.join(tmp, f"r{rank}.pt"))
    dist.destroy_process_group()


def test_gather_and_replay_world2(tmp_path):
    mp.spawn(_worker, args=(2, 29621, str(tmp_path)), nprocs=2, join=True)
    r0 = torch.load(tmp_path / "r0.pt")
    r1 = torch.load(tmp_path / "r1.pt")
    for name in r0["full"]:
        assert torch.equal(r0["full"][name], r1["full"][name]), name        # every rank holds the same batch
    assert r0["full"]["env"].tolist() == list(range(8))
    assert r0["full"]["reward"].tolist() == [1.0] * 4 + [2.0] * 4
    assert r0["full"]["done"].tolist() == [True, False, False, False, False, True, False, False]
    assert r0["full"]["block_bits"][:, 0].tolist() == [10] * 4 + [11] * 4
    assert r0["size"] == 6 and r0["head"] == 1
    assert r0["env"].tolist() == [7, 2, 4, 6, 5, 6]


def test_replay_sample_shapes():
    from bridges_b200.rollout import DeviceReplayBuffer, empty_batch
    replay = DeviceReplayBuffer(16, "cpu")
    b = empty_batch(5, "cpu")
    b["reward"] += torch.arange(5.0)
    replay.push(b)
    s = replay.sample(32, generator=torch.Generator().manual_seed(0))
    assert s["block_bits"].shape == (32, 64) and s["binary"].shape == (32, 6)
    assert set(s["reward"].tolist()) <= {0.0, 1.0, 2.0, 3.0, 4.0}


@pytest.mark.gpu
def test_lockstep_rollout_transitions_are_consistent():
    from bridges_b200.envs.batched import BatchedAssemblyGym
    from bridges_b200.rollout import DeviceReplayBuffer, random_policy, rollout_lockstep
    E = 64
    env = BatchedAssemblyGym(E, ["shapes/trapezoid.urdf"], max_steps=10)
    env.reset(dict(obstacles=[(0.6, 0, 0.3)], targets=[(0.6, 0, 0.9)]))
    replay = DeviceReplayBuffer(4096, env.device)
    xg = np.linspace(-2, 0, 10)
    for chunk in range(3):
        batch, keep = rollout_lockstep(env, random_policy(seed=chunk), 8, xg, replay=replay)
    assert len(replay) == 3 * 8 * E                      # every env always has a ground candidate
    data = {k: v[:len(replay)].cpu().numpy() for k, v in replay.data.items()}
    # the new raster is the old one plus the chosen candidate's raster, which never overlaps it
    assert np.array_equal(data["next_block_bits"], data["block_bits"] | data["action_bits"])
    assert not (data["block_bits"] & data["action_bits"]).any()
    assert (data["action_bits"] != 0).any(axis=1).all()
    # binary features: stable flag of the state before / after; unstable states end the episode with reward -1
    assert set(np.unique(data["binary"][:, 0])) <= {0.0, 1.0}
    unstable = data["next_binary"][:, 0] == 0
    assert data["done"][unstable].all() and (data["reward"][unstable] == -1).all()
    assert 0.05 < data["done"].mean() < 0.6
    # lin_reward is zero for unstable successors (successor_dqn.py:397-401)
    assert (data["lin_reward"][unstable] == 0).all()


@pytest.mark.gpu
def test_q_network_policy_batched_inference():
    """One batched pass of a Q-network with the reference's 5-argument signature over the valid candidates
    of all environments (successor_dqn.py:383-390 in lock-step form): the greedy choice equals a per-env
    argmax computed on the host, envs without candidates get a no-op, and a rollout driven by an
    nn.Module runs end to end."""
    from bridges_b200.envs.batched import BatchedAssemblyGym
    from bridges_b200.rollout import DeviceReplayBuffer, q_network_policy, rollout_lockstep
    E = 48
    env = BatchedAssemblyGym(E, ["shapes/trapezoid.urdf"], max_steps=10)
    env.reset(dict(obstacles=[(0.6, 0, 0.3)], targets=[(0.6, 0, 0.9)]))
    feats = env.observe(block=False, binary=False, obstacle=True, reward=True)
    xg = np.linspace(-2, 0, 10)
    seen = []

    def lin_q(block, binary, action, reward, obstacle):          # q = lin_reward of the candidate
        assert block.shape == action.shape == reward.shape == obstacle.shape and binary.shape == (block.shape[0], 6)
        seen.append(block.shape[0])
        return (action * reward).sum(dim=(1, 2, 3)), None, None

    # bring the envs to different states first
    rollout_lockstep(env, q_network_policy(lin_q, feats["reward"], feats["obstacle"], epsilon=1.0, seed=3), 2, xg,
                     gather=False)
    cand = env.enumerate_actions(xg, (0.0,), amax=128, with_bits=True)
    actions, index = q_network_policy(lin_q, feats["reward"], feats["obstacle"])(env, cand)
    env.sync()
    valid = cand["valid"].cpu().numpy().astype(bool)
    valid &= np.arange(128)[None, :] < cand["n"].cpu().numpy()[:, None]      # slots past n_cand are stale
    bits = cand["bits"].cpu().numpy().view(np.uint64)
    reward = feats["reward"].cpu().numpy()[:, 0]
    acts = actions.cpu().numpy().view(env.dt["action"])
    cands = cand["cand"].cpu().numpy().view(env.dt["action"]).reshape(E, 128)
    assert seen[-1] == int(valid.sum())                          # one row per valid candidate, nothing else
    for e in range(E):
        img = BatchedAssemblyGym.bits_to_bool(bits[e])           # [128, 64, 64]
        q = (img * reward[e][None]).sum(axis=(1, 2))
        q[~valid[e]] = -np.inf
        best = int(index[e].item())
        assert valid[e, best] and q[best] >= q.max() - 1e-5 * max(1.0, abs(q.max()))
        assert acts[e] == cands[e, best]

    class TinyNet(torch.nn.Module):                              # same interface as models/cv.py:76-105
        def __init__(self):
            super().__init__()
            self.lin = torch.nn.Linear(4 * 16 + 6, 1)

        def forward(self, block, binary, action, reward, obstacle):
            x = torch.cat([torch.nn.functional.adaptive_avg_pool2d(t, 4).flatten(1) for t in (block, action, reward, obstacle)], 1)
            return self.lin(torch.cat([x, binary], 1)).squeeze(1), None, None

    net = TinyNet().to(env.device)
    replay = DeviceReplayBuffer(2048, env.device)
    rollout_lockstep(env, q_network_policy(net, feats["reward"], feats["obstacle"], epsilon=0.2, seed=1), 6, xg,
                     replay=replay, gather=False)
    assert len(replay) == 6 * E
    data = {k: v[:len(replay)].cpu().numpy() for k, v in replay.data.items()}
    assert np.array_equal(data["next_block_bits"], data["block_bits"] | data["action_bits"])
    assert not (data["block_bits"] & data["action_bits"]).any()
