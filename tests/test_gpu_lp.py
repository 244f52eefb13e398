"""The warm-started LP verdict path of real steps (csrc/bw_lp.cuh) against the screen + Newton path it replaces.

Both paths answer `is_stable_rbe` (assembly_gym/utils/stability.py:49-71) for the two support variants of a step
(gym_env.py:235-245, 325-333); the oracle comparison of the LP path is tests/test_gpu_rollout_parity.py (it is the
default path).  Here the same action sequences run on two handles -- one created with the tuning hook BW_NO_LP -- and
every record is compared: the LP path must be a pure shortcut (identical verdicts outside the stated residual band,
identical everything else), it must actually be the path that decides, and its stored bases must survive what a
caller can do between steps (friction change, support change, resets with pre-placed blocks, refused actions)."""
import numpy as np
import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu

BAND = (1e-9, 1e-4)
XG = [-2.0 + 2.0 * i / 9 for i in range(10)]


def _tower(height, sq=0.6):
    return dict(obstacles=[(sq, 0, i * sq + sq / 2) for i in range(height - 1)], targets=[(sq, 0, (height - 1) * sq + sq / 2)])


def _bridge(n, sq=0.6):
    return dict(obstacles=[(i * sq, 0, sq / 2) for i in range(1, n + 1)], targets=[(n * sq + 2.5 * sq, 0, sq / 2)])


def _pair(monkeypatch, E, shapes, max_steps, share_h=None):
    """(LP handle, Newton-only handle) with the same configuration.  share_h: the shared-memory layout of
    multi-wave launches for the LP handle (tuning hook BW_SHARE_H: one packed matrix for both Newton problems, hence
    a smaller region for the LP, which then runs its two problems one after the other on one vector set)"""
    from bridges_b200.envs.batched import BatchedAssemblyGym
    urdfs = [f"shapes/{s}.urdf" for s in shapes]
    if share_h is not None:
        monkeypatch.setenv("BW_SHARE_H", str(share_h))
    a = BatchedAssemblyGym(E, urdfs, max_steps=max_steps)
    if share_h is not None:
        monkeypatch.delenv("BW_SHARE_H")
    monkeypatch.setenv("BW_NO_LP", "1")
    b = BatchedAssemblyGym(E, urdfs, max_steps=max_steps)
    monkeypatch.delenv("BW_NO_LP")
    return a, b


def _compare(oa, ob, tag, stats):
    """records of the LP handle (oa) and of the Newton handle (ob) after the same step"""
    for name in ("n_blocks", "n_interfaces", "reward", "truncated", "error", "n_targets_reached", "collision"):
        assert np.array_equal(oa[name], ob[name]), (tag, name)
    assert np.array_equal(oa["distance_to_targets"], ob["distance_to_targets"]), tag
    for verdict, res, bit in (("stable", "residual", 1), ("stable_unfrozen", "residual_unfrozen", 2)):
        # a verdict may differ only inside the residual band: where the Newton path's upper bound of r* lies in it,
        # or where the LP's basic solution shows r* <= its residual < 1e-4 (marginal piles with r* ~ 1e-8, which the
        # proximal iteration can give up on after its first stage; the oracle's BVLS residual sides with the LP there)
        r = ob[res]
        in_band = ((r > BAND[0]) & (r < BAND[1])) | ((oa[res] > BAND[0]) & (oa[res] < BAND[1]))
        undecided = ((oa["solver_status"] & bit) != 0) | ((ob["solver_status"] & bit) != 0)      # stable = None
        differs = oa[verdict] != ob[verdict]
        ok = in_band | undecided | ~differs
        assert ok.all(), (tag, verdict, np.nonzero(~ok)[0][:5], r[~ok][:5])
        stats["band"] += int((differs & ~undecided).sum())
        by_lp = (oa["solver_status"] & (16 * bit)) != 0
        stats["by_lp"] += int(by_lp.sum())
        stats["verdicts"] += len(r)
        # what the LP reports: the residual of its basic solution when stable (under the verdict threshold),
        # NaN with a certificate of infeasibility
        ra = oa[res]
        assert (ra[by_lp & (oa[verdict] == 1)] <= 1e-6).all(), (tag, res)
        assert np.isnan(ra[by_lp & (oa[verdict] == 0)]).all(), (tag, res)
    same = oa["stable"] == ob["stable"]
    assert np.array_equal(oa["terminated"][same], ob["terminated"][same]), tag
    stats["pivots"] += int(oa["lp_pivots"].sum())
    stats["newton_lp_handle"] += int(oa["newton_iters"].sum())
    stats["newton_plain"] += int(ob["newton_iters"].sum())


@pytest.mark.parametrize("case", ["tower2", "tower4_max15", "bridge5_mixed_max15", "bridge5_mixed_max15_lean_layout"])
def test_lp_path_is_a_pure_shortcut_of_the_newton_path(case, monkeypatch):
    cfg = {"tower2": (["trapezoid"], _tower(2), 10, 128),
           "tower4_max15": (["trapezoid"], _tower(4), 15, 256),
           "bridge5_mixed_max15": (["trapezoid", "hexagon"], _bridge(5), 15, 1024),
           "bridge5_mixed_max15_lean_layout": (["trapezoid", "hexagon"], _bridge(5), 15, 1024)}[case]
    shapes, task, max_steps, amax = cfg
    E, steps = 192, 70
    a, b = _pair(monkeypatch, E, shapes, max_steps, share_h=2 if case.endswith("lean_layout") else None)
    a.reset(task)
    b.reset(task)
    stats = dict(band=0, by_lp=0, verdicts=0, pivots=0, newton_lp_handle=0, newton_plain=0)
    # an in-band verdict may end an episode on one handle only: that environment is left out until both handles have
    # started a new episode (the Newton handle refuses the other's actions meanwhile and is reset every step)
    insync = np.ones(E, dtype=bool)
    compared = 0
    for k in range(steps):
        a.enumerate_actions(XG, (0.0,), amax=amax, with_bits=False)
        acts, _ = a.select_random(seed=4242 + 17 * k)
        a.step(acts)
        b.step(acts)
        oa, ob = a.read_out().copy(), b.read_out().copy()
        _compare(oa[insync], ob[insync], (case, k), stats)
        compared += int(insync.sum())
        insync &= (oa["terminated"] == ob["terminated"]) & (oa["stable"] == ob["stable"])
        a.reset_done()
        b.reset_done()
        na, nb_ = a.get_state()[1], b.get_state()[1]
        insync |= (na == 0) & (nb_ == 0)
        insync &= na == nb_
    assert compared > 0.97 * E * steps, compared
    assert stats["by_lp"] > 0.55 * stats["verdicts"], stats            # the LP path is the one that decides
    assert stats["band"] <= 0.002 * stats["verdicts"], stats          # verdicts that differ inside the residual band
    assert stats["newton_lp_handle"] < 0.05 * stats["newton_plain"], stats     # ... and the Newton solver is the exception
    assert 0.5 < stats["pivots"] / compared < 12, stats


def test_lp_bases_survive_friction_and_support_changes_and_prebuilt_resets(monkeypatch):
    """bw_set_mu / bw_set_static_mask / bw_reset(blocks=...) / a refused action between steps: the stored basis is
    dropped or extended as needed and the verdicts stay those of the Newton path."""
    from oracle import synth
    shapes = ["trapezoid", "hexagon", "cube1"]
    E = 96
    a, b = _pair(monkeypatch, E, shapes, None)
    rng = np.random.default_rng(77)
    lib = synth.library()
    plans = [synth.random_assembly(rng, lib, max_blocks=12, min_blocks=8) for _ in range(E)]
    stats = dict(band=0, by_lp=0, verdicts=0, pivots=0, newton_lp_handle=0, newton_plain=0)

    def act(p, k):
        return (p[k].target_block, p[k].target_face, p[k].shape, p[k].face, p[k].offset_x, p[k].offset_y) if k < len(p) else None

    def step_both(acts, tag):
        a.step(acts)
        b.step(acts)
        _compare(a.read_out().copy(), b.read_out().copy(), tag, stats)

    for env in (a, b):
        env.set_mu(0.8)
        env.reset(dict())
    for k in range(4):
        step_both([act(p, k) for p in plans], ("build", k))
    # friction drops in half of the environments: the stored bases were built with the old rays
    mus = np.where(np.arange(E) % 2 == 0, 0.8, 0.25)
    for env in (a, b):
        env.set_mu(mus)
    step_both([act(p, 4) for p in plans], ("after set_mu", 4))
    # an old block becomes a support in a third of the environments (AssemblyEnv.freeze_block): rows disappear
    blocks, n = a.get_state()
    mask = np.array([(1 << (int(n[e]) - 1)) | (1 if e % 3 == 0 else 0) for e in range(E)], dtype=np.uint32)
    for env in (a, b):
        env.set_static_mask(mask)
    step_both([act(p, 5) for p in plans], ("after set_static_mask", 5))
    # a refused action (face index out of range) leaves state and stored basis alone
    for env in (a, b):
        env.step([(0, 0, 0, 17, 0.0, 0.0)] * E)
        assert (env.read_out()["error"] == 1).all()
    step_both([act(p, 6) for p in plans], ("after a refused action", 6))
    step_both([act(p, 7) for p in plans], ("go on", 7))
    # reset with the current blocks pre-placed (reset(blocks=...), gym_env.py:255-289): the next step extends an
    # empty basis by every row at once
    blocks, n = a.get_state()
    tasks = [dict(blocks=[(blocks[e][i]["x"], blocks[e][i]["z"], blocks[e][i]["c"], blocks[e][i]["s"], blocks[e][i]["shape"])
                          for i in range(int(n[e]))]) for e in range(E)]
    for env in (a, b):
        env.reset(tasks)
    step_both([act(p, 8) if len(p) > 8 else None for p in plans], ("after a pre-built reset", 8))
    step_both([act(p, 9) if len(p) > 9 else None for p in plans], ("after a pre-built reset", 9))
    assert stats["by_lp"] > 0.3 * stats["verdicts"], stats
    assert stats["band"] <= 0.01 * stats["verdicts"], stats
