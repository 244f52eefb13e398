"""Fused rollout, packed-record replay ring and the transition gather.
CPU: ring bookkeeping and the chunk gather under gloo (world_size 2).  GPU: the records `bw_rollout_random` writes
replayed through the oracle, the unpack kernel, a Q-network policy around `bw_rollout_begin` / `bw_rollout_commit`."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
XG = [-2.0 + 2.0 * i / 9 for i in range(10)]


def _fake_records(n, env0, valid=None):
    from bridges_b200.rollout import record_dtype
    rec = np.zeros(n, dtype=record_dtype())
    rec["env"] = np.arange(n) + env0
    rec["reward"] = np.arange(n) + env0
    rec["valid"] = 1 if valid is None else valid
    rec["block_bits"][:, 0] = np.arange(n) + 100 * env0
    return torch.from_numpy(rec.view(np.uint8).reshape(n, -1).copy())


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    from bridges_b200.rollout import REC, TransitionRing, gather_records
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ring = TransitionRing(16, "cpu")
    for chunk in range(3):                         # 3 chunks of 4 records per rank -> 24 gathered records wrap a ring of 16
        mine = _fake_records(4, env0=10 * chunk + 4 * rank, valid=np.array([1, 1, 0, 1]) if rank == 1 else None)
        start, region = ring.reserve(world * 4)
        gather_records(mine, region)
    rec = ring.numpy()
    torch.save(dict(env=rec["env"].tolist(), valid=rec["valid"].tolist(), head=ring.head, size=ring.size, n=len(ring),
                    sampled=ring.column("env")[ring.sample_indices(64, torch.Generator().manual_seed(rank))].tolist()),
               os.path.join(tmp, f"r{rank}.pt"))
    dist.destroy_process_group()


def test_ring_and_gather_world2(tmp_path):
    mp.spawn(_worker, args=(2, 29631, str(tmp_path)), nprocs=2, join=True)
    r0 = torch.load(tmp_path / "r0.pt")
    r1 = torch.load(tmp_path / "r1.pt")
    assert r0["env"] == r1["env"] and r0["valid"] == r1["valid"]          # every rank holds the same ring
    # chunk 2 overwrote the first region; rank-major inside a chunk
    assert r0["env"] == [20, 21, 22, 23, 24, 25, 26, 27, 10, 11, 12, 13, 14, 15, 16, 17]
    assert r0["valid"] == [1, 1, 1, 1, 1, 1, 0, 1, 1, 1, 1, 1, 1, 1, 0, 1]
    assert r0["head"] == 8 and r0["size"] == 16 and r0["n"] == 14
    assert not {26, 16} & set(r0["sampled"]) and len(set(r0["sampled"])) > 8   # invalid records are never sampled


def test_ring_push_wraps_and_priorities():
    from bridges_b200.rollout import TransitionRing
    ring = TransitionRing(6, "cpu", prioritized=True)
    ring.push(_fake_records(4, 0), td_error=torch.tensor([0.0, 0.0, 5.0, 0.0]))
    ring.push(_fake_records(3, 4, valid=np.array([1, 0, 1])), td_error=torch.tensor([0.0, 9.0, 0.0]))
    rec = ring.numpy()
    assert rec["env"].tolist() == [6, 1, 2, 3, 4, 5] and ring.head == 1 and ring.size == 6 and len(ring) == 5
    idx = ring.sample_indices(4000, torch.Generator().manual_seed(0))
    env = ring.column("env")[idx]
    assert (env == 2).float().mean() > 0.99          # the record with the large TD error dominates; the invalid one never shows
    assert not (env == 5).any()
    ring.update_priorities(torch.tensor([1, 2]), torch.tensor([7.0, 0.0]))
    env = ring.column("env")[ring.sample_indices(4000, torch.Generator().manual_seed(1))]
    assert (env == 1).float().mean() > 0.99
    with pytest.raises(Exception):
        TransitionRing(10, "cpu").reserve(4)          # chunk size must divide the capacity


def _replay_records(rec, E, steps, cfg, obstacles, targets, watch=()):
    """Per environment the chain of records -> oracle replay job."""
    from tests import helpers as H
    by_env = {e: [None] * steps for e in range(E)}
    for r in rec:
        by_env[int(r["env"])][int(r["step"])] = r
    jobs = []
    for e in range(E):
        seq, reset_after = [], set()
        for k in range(steps):
            r = by_env[e][k]
            assert r is not None, (e, k)
            if not r["valid"]:
                seq.append(None)
                continue
            a = r["action"]
            seq.append((int(a["target_block"]), int(a["target_face"]), int(a["shape"]), int(a["face"]),
                        float(a["offset_x"]), float(a["offset_y"])))
            if r["done"] and not (r["terminated"] or r["truncated"]):
                reset_after.add(k)
        jobs.append(dict(shapes=cfg["shapes"], obstacles=obstacles, targets=targets, mu=0.8, max_steps=cfg["max_steps"],
                         actions=seq, x_ground=XG, offsets=(0.0,), reset_after=reset_after,
                         cand_steps=set(range(steps)) if e in watch else set()))
    return by_env, H.replay_parallel(jobs)


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["tower2", "bridge5_mixed_max15"])
def test_fused_rollout_records_match_oracle_replay(case):
    """The records written by `bw_rollout_random` (no torch op on the per-step path), replayed per environment
    through oracle.gym_env: state / action / next-state rasters, rewards, lin_reward, verdicts, binary features,
    episode ends incl. "no candidate left" (successor_dqn.py:403-411)."""
    from bridges_b200.envs.batched import BatchedAssemblyGym
    from bridges_b200.rollout import FusedRollout, TransitionRing
    from oracle import features as ofeat
    from tests import helpers as H
    from tests.test_gpu_rollout_parity import BAND, CASES
    cfg = CASES[case]
    E, T, chunks = 48, 8, 5
    steps = T * chunks
    obstacles, targets = cfg["task"]
    env = BatchedAssemblyGym(E, [H.URDF[n] for n in cfg["shapes"]], max_steps=cfg["max_steps"])
    env.reset(dict(obstacles=obstacles, targets=targets))
    ring = TransitionRing(steps * E, env.device)
    roll = FusedRollout(env, XG, (0.0,), amax=cfg["amax"], chunk_steps=T, ring=ring)
    assert roll.collect_random(chunks, seed=77) == steps
    roll.drain()
    env.sync()
    assert env.candidate_overflow() == 0
    rec = ring.numpy()
    assert len(rec) == steps * E and set(rec["step"]) == set(range(steps))
    watch = (0, 1, 2)
    by_env, traces = _replay_records(rec, E, steps, cfg, obstacles, targets, watch)
    oenv = H.oracle_env(cfg["shapes"], obstacles, targets)
    reward_f, _ = ofeat.get_task_features(oenv.reset()[0], H.XLIM, H.YLIM, H.IMG)
    n_checked = n_done = n_next = 0
    for e in range(E):
        prev_bits, prev_binary = [0] * 64, 1            # fresh environment: empty raster, stable
        for k in range(steps):
            r, ref, tag = by_env[e][k], traces[e][k], (case, e, k)
            if "cands" in ref and k > 0:
                p = by_env[e][k - 1]
                if p["valid"] and not p["done"]:        # the candidates of the next state counted by the rollout
                    assert int(p["n_next_candidates"]) == sum(ref["cand_mask"]), tag
                    n_next += 1
            if not r["valid"]:
                assert ref.get("skipped"), tag
                prev_bits, prev_binary = [0] * 64, 1
                continue
            assert [int(v) for v in r["block_bits"]] == prev_bits, tag                     # state before
            assert [int(v) for v in r["action_bits"]] == ref["new_bits"], tag              # chosen candidate's raster
            assert [int(v) for v in r["next_block_bits"]] == ref["bits"], tag              # state after
            assert int(r["binary"]) == prev_binary, tag
            in_band = ref["r_frozen"] is not None and BAND[0] < ref["r_frozen"] < BAND[1]
            in_band_u = ref["r_unfrozen"] is not None and BAND[0] < ref["r_unfrozen"] < BAND[1]
            if not in_band:
                assert bool(r["stable"]) == ref["frozen"] and float(r["reward"]) == ref["reward"], tag
                assert bool(r["terminated"]) == ref["terminated"] and bool(r["truncated"]) == ref["truncated"], tag
                assert int(r["next_binary"]) == int(ref["frozen"]), tag                    # collision flags are constant False
            if not in_band_u:
                assert bool(r["stable_unfrozen"]) == ref["stable_unfrozen"], tag
            if not in_band and not in_band_u:
                new_f = env.bits_to_bool(np.array(ref["new_bits"], dtype=np.uint64)).astype(np.float32)[None]
                want = float(ofeat.lin_reward(new_f, reward_f, ref["frozen"], ref["stable_unfrozen"]))
                assert abs(float(r["lin_reward"]) - want) <= 1e-5 * max(1.0, abs(want)), tag
            assert bool(r["done"]) >= bool(r["terminated"] or r["truncated"]), tag
            n_checked += 1
            n_done += bool(r["done"])
            if r["done"]:
                prev_bits, prev_binary = [0] * 64, 1
            else:
                prev_bits, prev_binary = ref["bits"], int(r["next_binary"])
    assert n_checked > 0.95 * E * steps and n_done > E and n_next > 50, (n_checked, n_done, n_next)


@pytest.mark.gpu
def test_unpack_kernel_and_sampling():
    from bridges_b200.envs.batched import BatchedAssemblyGym
    from bridges_b200.rollout import FusedRollout, TransitionRing
    E, T = 32, 4
    env = BatchedAssemblyGym(E, ["shapes/trapezoid.urdf"], max_steps=10)
    env.reset(dict(obstacles=[(0.6, 0, 0.3)], targets=[(0.6, 0, 0.9)]))
    ring = TransitionRing(4 * T * E, env.device, prioritized=True)
    roll = FusedRollout(env, XG, (0.0,), amax=128, chunk_steps=T, ring=ring)
    roll.collect_random(3, seed=5)
    roll.drain()
    batch = ring.sample(env, 96, generator=torch.Generator(device=env.device).manual_seed(0))
    env.sync()
    rec = ring.numpy()
    idx = batch["indices"].cpu().numpy()
    assert rec["valid"][idx].all()
    expand = lambda bits: BatchedAssemblyGym.bits_to_bool(bits).astype(np.float32)[:, None]
    assert np.array_equal(batch["block_features"].cpu().numpy(), expand(rec["block_bits"][idx]))
    assert np.array_equal(batch["action_features"].cpu().numpy(), expand(rec["action_bits"][idx]))
    assert np.array_equal(batch["next_block_features"].cpu().numpy(), expand(rec["next_block_bits"][idx]))
    bits6 = lambda b: ((b[:, None] >> np.arange(6)) & 1).astype(np.float32)
    assert np.array_equal(batch["binary_features"].cpu().numpy(), bits6(rec["binary"][idx]))
    assert np.array_equal(batch["next_binary_features"].cpu().numpy(), bits6(rec["next_binary"][idx]))
    assert np.array_equal(batch["reward"].cpu().numpy(), rec["reward"][idx])
    assert np.array_equal(batch["lin_reward"].cpu().numpy(), rec["lin_reward"][idx])
    assert np.array_equal(batch["done"].cpu().numpy(), rec["done"][idx].astype(bool))
    # next state = state + chosen candidate, which never overlaps it
    v = rec[rec["valid"] == 1]
    assert np.array_equal(v["next_block_bits"], v["block_bits"] | v["action_bits"]) and not (v["block_bits"] & v["action_bits"]).any()


@pytest.mark.gpu
def test_q_network_policy_around_begin_commit():
    """One batched pass of a Q-network with the reference's 5-argument signature over the valid candidates of all
    environments (successor_dqn.py:383-390 in lock-step form): the greedy choice equals a per-env argmax computed
    on the host, and a rollout driven by an nn.Module runs end to end through bw_rollout_commit."""
    from bridges_b200.envs.batched import BatchedAssemblyGym
    from bridges_b200.rollout import FusedRollout, TransitionRing, q_network_policy, rollout_policy
    E = 48
    env = BatchedAssemblyGym(E, ["shapes/trapezoid.urdf"], max_steps=10)
    env.reset(dict(obstacles=[(0.6, 0, 0.3)], targets=[(0.6, 0, 0.9)]))
    feats = env.observe(block=False, binary=False, obstacle=True, reward=True)
    roll = FusedRollout(env, XG, (0.0,), amax=128, chunk_steps=2)
    seen = []

    def lin_q(block, binary, action, reward, obstacle):          # q = lin_reward of the candidate
        assert block.shape == action.shape == reward.shape == obstacle.shape and binary.shape == (block.shape[0], 6)
        seen.append(block.shape[0])
        return (action * reward).sum(dim=(1, 2, 3)), None, None

    rollout_policy(roll, q_network_policy(lin_q, feats["reward"], feats["obstacle"], epsilon=1.0, seed=3), 2)
    cand = roll.candidates()
    index = q_network_policy(lin_q, feats["reward"], feats["obstacle"])(env, cand)
    env.sync()
    valid = cand["valid"].cpu().numpy().astype(bool)
    valid &= np.arange(128)[None, :] < cand["n"].cpu().numpy()[:, None]
    assert np.array_equal(valid.sum(axis=1), cand["n_valid"].cpu().numpy())
    bits = cand["bits"].dense() if hasattr(cand["bits"], "dense") else cand["bits"]     # store or dense copies
    bits = bits.cpu().numpy().view(np.uint64)
    reward = feats["reward"].cpu().numpy()[:, 0]
    cands = cand["cand"].cpu().numpy().view(env.dt["action"]).reshape(E, 128)
    assert seen[-1] == int(valid.sum())
    for e in range(E):
        img = BatchedAssemblyGym.bits_to_bool(bits[e])
        q = (img * reward[e][None]).sum(axis=(1, 2))
        q[~valid[e]] = -np.inf
        best = int(index[e].item())
        assert valid[e, best] and q[best] >= q.max() - 1e-5 * max(1.0, abs(q.max()))
    chosen = cands[np.arange(E), index.cpu().numpy()].copy()
    rec = roll.step(index).cpu().numpy().reshape(-1).view(env.dt["transition"]).copy()
    assert rec["valid"].all() and (rec["action"] == chosen).all()

    class TinyNet(torch.nn.Module):                              # same interface as models/cv.py:76-105
        def __init__(self):
            super().__init__()
            self.lin = torch.nn.Linear(4 * 16 + 6, 1)

        def forward(self, block, binary, action, reward, obstacle):
            x = torch.cat([torch.nn.functional.adaptive_avg_pool2d(t, 4).flatten(1) for t in (block, action, reward, obstacle)], 1)
            return self.lin(torch.cat([x, binary], 1)).squeeze(1), None, None

    net = TinyNet().to(env.device)
    ring = TransitionRing(2048, env.device)
    rollout_policy(roll, q_network_policy(net, feats["reward"], feats["obstacle"], epsilon=0.2, seed=1), 6, ring=ring)
    env.sync()
    assert len(ring) == 6 * E
    v = ring.numpy()
    assert np.array_equal(v["next_block_bits"], v["block_bits"] | v["action_bits"])


@pytest.mark.gpu
def test_fused_rollout_with_and_without_candidate_store(monkeypatch):
    """With a candidate store an iteration of `bw_rollout_random` is two launches (step; the candidate kernel records
    the step, restarts finished episodes, closes the iteration and picks for the next one), without one
    (BW_CAND_CACHE_MB=0: plain candidate kernel, dense raster copies) the stand-alone pick / record / finalize kernels
    run.  Same seeds, same records, byte for byte -- also across calls (the first iteration of a call has its own pick
    kernel) and around a caller-driven `step` in between."""
    import torch
    from bridges_b200.envs.batched import BatchedAssemblyGym
    from bridges_b200.rollout import FusedRollout, TransitionRing
    from tests import helpers as H
    from tests.test_gpu_rollout_parity import CASES
    cfg = CASES["bridge5_mixed_max15"]
    E, T = 64, 6
    obstacles, targets = cfg["task"]

    def make():
        env = BatchedAssemblyGym(E, [H.URDF[n] for n in cfg["shapes"]], max_steps=cfg["max_steps"])
        env.reset(dict(obstacles=obstacles, targets=targets))
        ring = TransitionRing(5 * T * E, env.device)
        return env, ring, FusedRollout(env, XG, (0.0,), amax=cfg["amax"], chunk_steps=T, ring=ring)

    a_env, a_ring, a_roll = make()
    monkeypatch.setenv("BW_CAND_CACHE_MB", "0")
    b_env, b_ring, b_roll = make()
    monkeypatch.delenv("BW_CAND_CACHE_MB")
    assert a_roll.candidates()["slot"] is not None and b_roll.candidates()["slot"] is None
    for roll in (a_roll, b_roll):
        roll.collect_random(2, seed=5)
    # T iterations around a caller's choice: the first valid candidate of every environment
    for _ in range(T):
        idx = []
        for roll in (a_roll, b_roll):
            c = roll.candidates()
            valid = c["valid"].bool() & (torch.arange(c["amax"], device=c["valid"].device)[None, :] < c["n"][:, None])
            index = torch.where(valid.any(dim=1), valid.float().argmax(dim=1), torch.full((E,), -1, device=valid.device)).to(torch.int32)
            idx.append(index.cpu().numpy())
            roll.ring.push(roll.step(index))
        assert np.array_equal(idx[0], idx[1])
    for roll in (a_roll, b_roll):
        roll.collect_random(2, seed=9)
        roll.drain()
        roll.env.sync()
    ra, rb = a_ring.numpy(), b_ring.numpy()
    assert len(ra) == len(rb) == 5 * T * E
    valid = ra["valid"] == 1
    assert np.array_equal(ra["valid"], rb["valid"]) and valid.sum() > 0.9 * len(ra)
    for name in ra.dtype.names:
        if name == "reserved":
            continue
        assert np.array_equal(ra[name][valid], rb[name][valid]), name
    assert np.array_equal(ra["done"], rb["done"]) and np.array_equal(ra["env"], rb["env"]) and np.array_equal(ra["step"], rb["step"])
    assert (ra["done"][valid] == 1).sum() > E              # episodes ended and were restarted on both paths
    a_env.close(); b_env.close()
