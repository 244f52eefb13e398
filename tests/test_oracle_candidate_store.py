"""The facts about the REFERENCE's candidate stage that the CUDA candidate store relies on (DESIGN.md section 6,
`enumerate_store_kernel`), checked on the oracle (robotoddler/utils/actions.py:7-82, successor_dqn.py:88-94,
gym_env.py:304-323 restated):

1. within an episode the block raster only gains pixels: raster(t + 1) = raster(t) | raster(new block);
2. a candidate is identified by (shape, face, target block, target face, offset): as long as it stays listed, its
   feature raster and its `collision_on_action` flag do not change from step to step (they depend on the block library
   and on the pose of the target block only);
3. hence a candidate that `filter_actions` has dropped for overlap stays dropped until the episode ends, and a kept one
   can only be dropped by the pixels the raster gained.
"""
import numpy as np
import pytest

from oracle import actions as oact
from oracle import features as ofeat
from tests import helpers as H

XG = np.linspace(-2, 0, 10)
CASES = {
    "tower2": dict(shapes=["trapezoid"], obstacles=[], targets=[(0.0, 0.0, 2.5)], offsets=(0.0,)),
    "bridge_mixed": dict(shapes=["trapezoid", "hexagon"], obstacles=[(0.6, 0, 0.3), (1.2, 0, 0.3)], targets=[(2.4, 0, 0.3)],
                         offsets=(0.0, 0.25)),
}


def _key(a):
    return (a.shape, a.face, a.target_block, a.target_face, float(a.offset_x))


@pytest.mark.parametrize("case", sorted(CASES))
def test_candidates_keep_raster_and_bounds_flag_and_rasters_only_grow(case):
    cfg = CASES[case]
    rng = np.random.default_rng(11)
    env = H.oracle_env(cfg["shapes"], cfg["obstacles"], cfg["targets"], max_steps=6)
    n_kept = n_episodes = 0
    for episode in range(2):
        obs, _ = env.reset()
        _, obstacle_f = ofeat.get_task_features(obs, H.XLIM, H.YLIM, H.IMG)
        seen, prev_raster = {}, np.zeros(H.IMG, dtype=bool)
        for step in range(5):
            block_f, _ = ofeat.get_state_features(obs, H.XLIM, H.YLIM, H.IMG)
            raster = block_f[0] > 0
            assert not (prev_raster & ~raster).any()                      # 1. nothing is ever erased
            cands = list(oact.generate_actions(env, XG, cfg["offsets"]))
            feats = ofeat.get_action_features(env, cands, H.XLIM, H.YLIM, H.IMG)
            _, _, mask = oact.filter_actions(env, cands, feats, block_f, obstacle_f, H.XLIM, H.YLIM)
            for a, f, ok in zip(cands, feats, mask):
                bad = bool(env.collision_on_action(a, H.XLIM, H.YLIM))
                img = f[0] > 0
                if _key(a) in seen:
                    old_img, old_bad, old_ok = seen[_key(a)]
                    assert np.array_equal(img, old_img) and bad == old_bad      # 2. same raster, same bounds flag
                    assert not (ok and not old_ok)                               # 3. dropped stays dropped
                    if old_ok and not ok:                                        # ... and only new pixels drop one
                        assert (img & raster & ~prev_raster).any()
                    n_kept += 1
                seen[_key(a)] = (img, bad, bool(ok))
            valid = np.flatnonzero(mask)
            if valid.size == 0:
                break
            a = cands[int(rng.choice(valid))]
            new_img = feats[cands.index(a)][0] > 0
            obs, _, terminated, truncated, _ = env.step(a)
            prev_raster = raster
            after = ofeat.get_state_features(obs, H.XLIM, H.YLIM, H.IMG)[0][0] > 0
            assert np.array_equal(after, raster | new_img)                        # 1. ... and what is gained is the new block
            if terminated or truncated:
                break
        n_episodes += 1
    assert n_episodes == 2 and n_kept > 50


@pytest.mark.parametrize("case", sorted(CASES))
def test_store_restatement_equals_recomputing_from_scratch(case):
    """oracle/candidate_store.py (the rules of `enumerate_store_kernel`: slots, dropped blocks, fresh calls, stamps,
    new-pixels-only tests) against generate_actions + get_action_features + filter_actions recomputed at every call:
    same candidates, same rasters, same validity -- through episode ends, calls the store skips (two blocks of new
    pixels), repeated calls on one state and a store call in the middle of an episode it has not seen."""
    from oracle.candidate_store import CandidateStore
    cfg = CASES[case]
    rng = np.random.default_rng(5)
    env = H.oracle_env(cfg["shapes"], cfg["obstacles"], cfg["targets"], max_steps=7)
    store = CandidateStore(XG, cfg["offsets"], H.XLIM, H.YLIM, H.IMG)
    calls = 0
    for episode in range(4):
        obs, _ = env.reset()
        _, obstacle_f = ofeat.get_task_features(obs, H.XLIM, H.YLIM, H.IMG)
        for step in range(6):
            block_f, _ = ofeat.get_state_features(obs, H.XLIM, H.YLIM, H.IMG)
            cands = list(oact.generate_actions(env, XG, cfg["offsets"]))
            feats = ofeat.get_action_features(env, cands, H.XLIM, H.YLIM, H.IMG)
            _, _, mask = oact.filter_actions(env, cands, feats, block_f, obstacle_f, H.XLIM, H.YLIM)
            skip = (episode * 6 + step) % 5 == 3 or (episode == 2 and step < 2)      # the store does not see every state
            if not skip:
                for _ in range(2 if step == 2 else 1):                                # ... and sees some of them twice
                    got_a, got_mask, got_img = store.enumerate(env, block_f[0] > 0, obstacle_f[0] > 0)
                    calls += 1
                    assert [_key(a) for a in got_a] == [_key(a) for a in cands]
                    assert np.array_equal(got_mask, mask), (case, episode, step, np.flatnonzero(got_mask != mask))
                    assert all(np.array_equal(g, f[0] > 0) for g, f in zip(got_img, feats))
            valid = np.flatnonzero(mask)
            if valid.size == 0:
                break
            # every episode opens with the same block: its slots outlive the reset with the verdicts and the stamps of
            # the previous episode, and must not be believed (a slot is only trusted when the PREVIOUS call listed it)
            choice = valid[0] if step == 0 else rng.choice(valid)
            if step == 1:
                # ... the second block of the even episodes is the one whose raster covers most of the other valid
                # candidates (their slots end the episode with the verdict "overlaps"), the odd episodes take the one
                # that covers fewest: the same slots are listed again there and are valid
                imgs = [feats[i][0] > 0 for i in valid]
                cover = [sum(bool((imgs[j] & imgs[k]).any()) for k in range(len(valid)) if k != j) for j in range(len(valid))]
                choice = valid[int(np.argmax(cover) if episode % 2 == 0 else np.argmin(cover))]
            obs, _, terminated, truncated, _ = env.step(cands[int(choice)])
            if terminated or truncated:
                break
    st = store.stats
    assert calls > 10 and st["posed"] > 50 and st["incremental_tests"] > 50 and st["no_test"] > 20 and st["full_tests"] > 20, dict(st)
    # the point of the store: most listed candidates are neither posed nor tested in full
    assert st["incremental_tests"] + st["no_test"] > st["posed"], dict(st)
