"""Parity of the candidate-action kernel, task features and lin_reward with the oracle."""
import numpy as np
import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu

XG = np.linspace(-2, 0, 10)


def _run(shape_names, urdfs, obstacles, targets, actions, mu=0.8, shape_kwargs=None, offsets=(0.0,)):
    import torch
    from bridges_b200.envs.batched import BatchedAssemblyGym
    from oracle import actions as oact
    from oracle import features as ofeat
    from oracle.gym_env import Action as OAction
    oenv = H.oracle_env(shape_names, obstacles, targets, mu=mu, shape_kwargs=shape_kwargs)
    obs, _ = oenv.reset()
    reward_f, obstacle_f = ofeat.get_task_features(obs, H.XLIM, H.YLIM, H.IMG)
    env = BatchedAssemblyGym(2, urdfs, mu=mu)
    if shape_kwargs:
        from bridges_b200.envs.assembly_env import Shape
        env.set_shapes([Shape(urdf_file=H.URDF[n], **shape_kwargs.get(i, {})) for i, n in enumerate(shape_names)])
    env.reset(dict(obstacles=obstacles, targets=targets))
    feats = env.observe(block=True, binary=True, obstacle=True, reward=True)
    assert np.array_equal(feats["obstacle"][0].cpu().numpy(), obstacle_f)                 # bit-exact
    assert np.allclose(feats["reward"][0].cpu().numpy(), reward_f, rtol=1e-5, atol=1e-7)  # f32 convolution
    for k, a in enumerate([None] + list(actions)):
        if a is not None:
            obs, reward, terminated, truncated, _ = oenv.step(OAction(*a))
            frozen, unfrozen = oenv.stabilities_freezing()
            env.step([a, a])
            out = env.read_out()
            new_block = oenv.assembly_env.blocks[-1]
            from oracle.rendering import render_blocks_2d
            action_f = render_blocks_2d([new_block], H.XLIM, H.YLIM, H.IMG).astype(np.float32)[None]
            want = ofeat.lin_reward(action_f, reward_f, frozen, unfrozen)
            assert abs(float(out[0]["lin_reward"]) - float(want)) <= 1e-5 * max(1.0, abs(float(want))), k
        block_f, binary_f = ofeat.get_state_features(obs, H.XLIM, H.YLIM, H.IMG)
        feats = env.observe()
        assert np.array_equal(feats["block"][1].cpu().numpy(), block_f), k
        assert np.array_equal(feats["binary"][1].cpu().numpy(), binary_f), k
        cands = [*oact.generate_actions(oenv, XG, list(offsets))]
        cand_f = ofeat.get_action_features(oenv, cands, H.XLIM, H.YLIM, H.IMG)
        _, _, mask = oact.filter_actions(oenv, cands, cand_f, block_f, obstacle_f, H.XLIM, H.YLIM)
        c = env.enumerate_actions(XG, offsets, amax=256)
        env.sync()
        n = int(c["n"][0].item())
        assert n == len(cands), (k, n, len(cands))
        got = c["cand"].cpu().numpy().view(env.dt["action"]).reshape(2, 256)[0][:n]
        for i, ca in enumerate(cands):
            assert (got[i]["target_block"], got[i]["target_face"], got[i]["shape"], got[i]["face"],
                    got[i]["offset_x"], got[i]["offset_y"]) == \
                   (ca.target_block, ca.target_face, ca.shape, ca.face, ca.offset_x, ca.offset_y), (k, i)
        assert np.array_equal(c["valid"][0, :n].cpu().numpy().astype(bool), mask), k
        img = env.expand_bits(c["bits"][0, :n].contiguous())
        assert np.array_equal(img.cpu().numpy(), cand_f), k                                # bit-exact rasters
    return env




def test_bridge_golden_candidates():
    obstacles = [(i * 0.6, 0, 0.3) for i in range(1, 8)]
    targets = [(7 * 0.6 + 2.5 * 0.6, 0, 0.3)]
    actions = [(-1, 0, 0, 2, -0.45), (0, 0, 0, 1, 0), (1, 3, 0, 1, 0), (2, 3, 0, 0, 0), (3, 3, 0, 1, 0),
               (4, 3, 0, 1, 0), (5, 3, 0, 3, 0), (6, 1, 0, 2, 0)]
    _run(["trapezoid"], [H.URDF["trapezoid"]], obstacles, targets, actions, mu=2.0)


def test_mixed_library_with_face_restrictions_and_offsets():
    actions = [(-1, 0, 1, 0, 0, 0), (-1, 0, 1, 0, -1.2, 0), (0, 3, 0, 3, 0.25, 0), (2, 1, 1, 0, 0, 0)]
    _run(["trapezoid", "cube1"], [H.URDF["trapezoid"], H.URDF["cube1"]], [[0, 0, 2.0]],
         [[0, 0, 0.5], [0, 0, 5.5]], actions,
         shape_kwargs={1: dict(receiving_faces_2d=[0], target_faces_2d=[2])}, offsets=(0.0, 0.25, -0.25))


def test_hexagon_candidates_and_random_policy():
    import torch
    env = _run(["hexagon"], [H.URDF["hexagon"]], [(0.6, 0, 0.3)], [(0.6, 0, 0.9)],
               [(-1, 0, 0, 0, -1.0, 0), (0, 5, 0, 0, 0, 0)])
    # the synthetic policy only ever picks valid candidates and is reproducible
    c = env.enumerate_actions(XG, (0.0,), amax=256)
    a1, i1 = env.select_random(seed=5)
    i1 = i1.clone()
    a2, i2 = env.select_random(seed=5)
    env.sync()
    assert torch.equal(i1, i2)
    valid = c["valid"].cpu().numpy()
    for e, idx in enumerate(i1.cpu().numpy()):
        assert idx >= 0 and valid[e, idx] == 1


def test_candidate_cache_matches_plain_kernel(monkeypatch):
    """bw_enumerate_actions keeps candidate placements and rasters between calls (enumerate_kernel<true>); a handle
    created with BW_CAND_CACHE_MB=0 runs the plain kernel.  Same rollouts, same outputs: candidates, validity and
    rasters of every call -- through auto-resets, a change of the offset table and a reset with pre-placed blocks."""
    import torch
    from bridges_b200.envs.batched import BatchedAssemblyGym
    E, amax = 48, 512
    urdfs = [H.URDF["trapezoid"], H.URDF["hexagon"]]
    task = dict(obstacles=[(0.6, 0, 0.3), (1.2, 0, 0.3)], targets=[(2.4, 0, 0.3)])
    cached = BatchedAssemblyGym(E, urdfs, max_steps=10)
    monkeypatch.setenv("BW_CAND_CACHE_MB", "0")
    plain = BatchedAssemblyGym(E, urdfs, max_steps=10)
    monkeypatch.delenv("BW_CAND_CACHE_MB")
    for env in (cached, plain):
        env.reset(task)

    def compare(offsets, tag):
        a = cached.enumerate_actions(XG, offsets, amax=amax)
        b = plain.enumerate_actions(XG, offsets, amax=amax)
        cached.sync(); plain.sync()
        n = a["n"].cpu().numpy()
        assert np.array_equal(n, b["n"].cpu().numpy()), tag
        sz = cached.dt["action"].itemsize
        ca = a["cand"].cpu().numpy().reshape(E, amax, sz)
        cb = b["cand"].cpu().numpy().reshape(E, amax, sz)
        va, vb = a["valid"].cpu().numpy(), b["valid"].cpu().numpy()
        ba, bb = a["bits"].cpu().numpy(), b["bits"].cpu().numpy()
        for e in range(E):
            assert np.array_equal(ca[e, :n[e]], cb[e, :n[e]]), (tag, e)
            assert np.array_equal(va[e, :n[e]], vb[e, :n[e]]), (tag, e)
            assert np.array_equal(ba[e, :n[e]], bb[e, :n[e]]), (tag, e)
        return int(n.max()), int(va.sum())

    n_max = n_valid = 0
    for k in range(36):
        offsets = (0.0,) if k < 24 else (0.0, 0.25)          # the offset table changes: every slot is stale
        m, v = compare(offsets, k)
        n_max, n_valid = max(n_max, m), n_valid + v
        acts, _ = cached.select_random(seed=1000 + k)
        acts = acts.clone()
        cached.step(acts); plain.step(acts)
        if k == 17:                                           # reset with pre-placed blocks, all environments
            pre = dict(task, blocks=[(-1.5, 0.5, 1.0, 0.0, 1), (3.5, 0.4, 0.0, 1.0, 0)])     # (x, z, c, s, shape)
            cached.reset(pre); plain.reset(pre)
        else:
            cached.reset_done(); plain.reset_done()
    assert n_max > 150 and n_valid > 5000
    oa, ob = cached.read_out(), plain.read_out()
    assert np.array_equal(oa["n_blocks"], ob["n_blocks"])
    cached.close(); plain.close()


def test_stored_candidates_match_plain_kernel(monkeypatch):
    """bw_enumerate_actions_stored leaves the rasters in the handle's candidate store and tests a candidate that the
    previous call listed against the NEW pixels of the block raster only (enumerate_store_kernel<false>).  Same
    rollouts as a handle without a store (plain kernel, dense copies): candidates, validity and the gathered rasters
    of every call -- through auto-resets, calls that are skipped (two blocks of new pixels), repeated calls on one
    state, a dense call in between, a change of the offset table and a reset with pre-placed blocks."""
    import torch
    from bridges_b200.envs.batched import BatchedAssemblyGym
    E, amax = 48, 512
    urdfs = [H.URDF["trapezoid"], H.URDF["hexagon"]]
    task = dict(obstacles=[(0.6, 0, 0.3), (1.2, 0, 0.3)], targets=[(2.4, 0, 0.3)])
    stored = BatchedAssemblyGym(E, urdfs, max_steps=10)
    monkeypatch.setenv("BW_CAND_CACHE_MB", "0")
    plain = BatchedAssemblyGym(E, urdfs, max_steps=10)
    monkeypatch.delenv("BW_CAND_CACHE_MB")
    for env in (stored, plain):
        env.reset(task)
    sz = stored.dt["action"].itemsize

    def compare(offsets, tag, mode="stored"):
        a = stored.enumerate_actions(XG, offsets, amax=amax, with_bits=mode)
        b = plain.enumerate_actions(XG, offsets, amax=amax)
        ba = a["bits"].dense() if mode == "stored" else a["bits"]
        stored.sync(); plain.sync()
        n = a["n"].cpu().numpy()
        assert np.array_equal(n, b["n"].cpu().numpy()), tag
        ca = a["cand"].cpu().numpy().reshape(E, amax, sz)
        cb = b["cand"].cpu().numpy().reshape(E, amax, sz)
        va, vb = a["valid"].cpu().numpy(), b["valid"].cpu().numpy()
        ba, bb = ba.cpu().numpy(), b["bits"].cpu().numpy()
        for e in range(E):
            assert np.array_equal(ca[e, :n[e]], cb[e, :n[e]]), (tag, e)
            assert np.array_equal(va[e, :n[e]], vb[e, :n[e]]), (tag, e, np.flatnonzero(va[e, :n[e]] != vb[e, :n[e]]))
            assert np.array_equal(ba[e, :n[e]], bb[e, :n[e]]), (tag, e)
        if mode == "stored":      # a gather of chosen candidates: the first valid one of every environment
            first = torch.as_tensor(np.argmax(vb, axis=1).astype(np.int32))
            got = a["bits"][torch.arange(E), first].cpu().numpy()
            assert np.array_equal(got, bb[np.arange(E), first.numpy()]), tag
        return int(n.max()), int(va.sum()), b

    n_max = n_valid = 0
    for k in range(60):
        offsets = (0.0,) if k < 40 else (0.0, 0.25)          # the offset table changes: every slot is stale
        if k % 7 == 3:                                        # the stored side skips this call
            b = plain.enumerate_actions(XG, offsets, amax=amax)
        else:
            m, v, b = compare(offsets, k, "stored" if k % 11 != 5 else True)
            n_max, n_valid = max(n_max, m), n_valid + v
            if k % 5 == 2:                                    # once more on the same state: nothing new to test
                compare(offsets, (k, "again"))
        acts, _ = plain.select_random(seed=2000 + k, cand=b)
        acts = acts.clone()
        stored.select_random(seed=2000 + k, cand=b)           # same states, same choice: the "no candidate" flags agree
        stored.step(acts); plain.step(acts)
        if k == 23:                                           # reset with pre-placed blocks, all environments
            pre = dict(task, blocks=[(-1.5, 0.5, 1.0, 0.0, 1), (3.5, 0.4, 0.0, 1.0, 0)])     # (x, z, c, s, shape)
            stored.reset(pre); plain.reset(pre)
        else:
            stored.reset_done(); plain.reset_done()
    assert n_max > 150 and n_valid > 5000
    oa, ob = stored.read_out(), plain.read_out()
    assert np.array_equal(oa["n_blocks"], ob["n_blocks"])
    # a handle without a store answers with_bits="stored" with dense copies behind the same interface
    d = plain.enumerate_actions(XG, (0.0, 0.25), amax=amax, with_bits="stored")
    s2 = stored.enumerate_actions(XG, (0.0, 0.25), amax=amax, with_bits="stored")
    plain.sync(); stored.sync()
    n = d["n"].cpu().numpy()
    da, sa = d["bits"].dense().cpu().numpy(), s2["bits"].dense().cpu().numpy()
    assert all(np.array_equal(da[e, :n[e]], sa[e, :n[e]]) for e in range(E))
    er, ar = torch.arange(E), torch.zeros(E, dtype=torch.int64)
    assert np.array_equal(d["bits"][er, ar].cpu().numpy(), s2["bits"][er, ar].cpu().numpy())
    stored.close(); plain.close()


def test_stored_verdicts_are_not_believed_across_episodes(monkeypatch):
    """The scenario that separates "listed by the previous call" from "has a verdict" (tests/test_oracle_candidate_store.py
    holds the same rules against the oracle on the CPU): every episode opens with the same block, so its candidates' slots
    outlive the reset; the second block of the even episodes covers as many of them as possible (they end the episode
    with the verdict "overlaps"), the odd episodes place their second block where it covers fewest -- the same slots are
    listed again and are valid.  Stored candidates against the plain kernel, call by call."""
    from bridges_b200.envs.batched import BatchedAssemblyGym
    E, amax = 4, 512
    urdfs = [H.URDF["trapezoid"], H.URDF["hexagon"]]
    task = dict(obstacles=[(0.6, 0, 0.3), (1.2, 0, 0.3)], targets=[(2.4, 0, 0.3)])
    stored = BatchedAssemblyGym(E, urdfs, max_steps=6)
    monkeypatch.setenv("BW_CAND_CACHE_MB", "0")
    plain = BatchedAssemblyGym(E, urdfs, max_steps=6)
    monkeypatch.delenv("BW_CAND_CACHE_MB")
    offsets = (0.0, 0.25)
    sz = stored.dt["action"].itemsize
    dropped_then_valid = 0
    end_of_even = None                    # validity of the first block's candidates at the end of an even episode
    for episode in range(4):
        for env in (stored, plain):
            env.reset(task)
        for step in range(4):
            a = stored.enumerate_actions(XG, offsets, amax=amax, with_bits="stored")
            b = plain.enumerate_actions(XG, offsets, amax=amax)
            stored.sync(); plain.sync()
            n = int(b["n"][0].item())
            va, vb = a["valid"].cpu().numpy()[:, :n], b["valid"].cpu().numpy()[:, :n]
            assert np.array_equal(a["n"].cpu().numpy(), b["n"].cpu().numpy())
            assert np.array_equal(va, vb), (episode, step, np.flatnonzero(va[0] != vb[0]))
            assert np.array_equal(a["bits"].dense().cpu().numpy()[:, :n], b["bits"].cpu().numpy()[:, :n])
            cands = b["cand"].cpu().numpy().reshape(E, amax, sz)[0, :n].copy().view(stored.dt["action"]).reshape(n)
            valid = np.flatnonzero(vb[0])
            assert valid.size > 0
            if step == 1:                 # one block placed: the candidates of that block, keyed by what identifies them
                keys = [(int(c["shape"]), int(c["face"]), int(c["target_block"]), int(c["target_face"]), float(c["offset_x"])) for c in cands]
                now = {k: bool(v) for k, v in zip(keys, vb[0]) if k[2] == 0}
                if episode % 2 == 1 and end_of_even is not None:
                    dropped_then_valid += sum(1 for k, ok in now.items() if ok and end_of_even.get(k) is False)
            choice = int(valid[0])
            if step == 1:
                img = BatchedAssemblyGym.bits_to_bool(b["bits"][0, :n].cpu().numpy().view(np.uint64))[valid]
                cover = [int((img[j][None] & img).any(axis=(1, 2)).sum()) for j in range(len(valid))]
                choice = int(valid[int(np.argmax(cover) if episode % 2 == 0 else np.argmin(cover))])
            elif step > 1:
                choice = int(valid[(7 * episode + 3 * step) % valid.size])
            acts = np.zeros(E, dtype=stored.dt["action"])
            acts[:] = cands[choice]
            stored.step(acts); plain.step(acts)
            if step == 1 and episode % 2 == 0:
                c = plain.enumerate_actions(XG, offsets, amax=amax)
                plain.sync()
                m = int(c["n"][0].item())
                cc = c["cand"].cpu().numpy().reshape(E, amax, sz)[0, :m].copy().view(stored.dt["action"]).reshape(m)
                end_of_even = {(int(q["shape"]), int(q["face"]), int(q["target_block"]), int(q["target_face"]), float(q["offset_x"])): bool(v)
                               for q, v in zip(cc, c["valid"].cpu().numpy()[0, :m]) if int(q["target_block"]) == 0}
    assert dropped_then_valid > 0         # the scenario happened: slots with the verdict "overlaps" came back valid
    oa, ob = stored.read_out(), plain.read_out()
    assert np.array_equal(oa["n_blocks"], ob["n_blocks"])
    stored.close(); plain.close()
