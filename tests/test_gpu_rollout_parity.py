"""The path the benchmark times -- enumerate -> select_random -> step -> reset_done over many auto-resets --
replayed environment by environment through the CPU oracle (oracle.gym_env, gym_env.py:218-253,325-333).

This is the only test in which verdicts are carried from step to step on the GPU (`prev_released_ok`, the
released-block verdict of step t standing in for the frozen solve of step t + 1, csrc/bw_step.cu), so every
record of every step is compared: poses, rasters, interfaces, verdicts outside the residual band, rewards,
termination, targets, distances, lin_reward, and on two environments per run the candidate list and its
validity mask as well."""
import numpy as np
import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu

BAND = (1e-9, 1e-4)
XG = [-2.0 + 2.0 * i / 9 for i in range(10)]           # np.linspace(-2, 0, 10), successor_dqn.py:611


def _tower(height, sq=0.6):
    return ([(sq, 0, i * sq + sq / 2) for i in range(height - 1)], [(sq, 0, (height - 1) * sq + sq / 2)])


def _bridge(n, sq=0.6):
    return ([(i * sq, 0, sq / 2) for i in range(1, n + 1)], [(n * sq + 2.5 * sq, 0, sq / 2)])


CASES = {
    "tower2": dict(shapes=["trapezoid"], task=_tower(2), max_steps=10, amax=128, steps=48),
    "tower4_max15": dict(shapes=["trapezoid"], task=_tower(4), max_steps=15, amax=256, steps=56),
    "bridge5_mixed_max15": dict(shapes=["trapezoid", "hexagon"], task=_bridge(5), max_steps=15, amax=1024, steps=40),
}


@pytest.mark.parametrize("case", list(CASES))
def test_lockstep_autoreset_rollout_matches_oracle_replay(case):
    import torch
    from bridges_b200.envs.batched import BatchedAssemblyGym
    from oracle import features as ofeat
    cfg = CASES[case]
    E, steps, amax = 64, cfg["steps"], cfg["amax"]
    obstacles, targets = cfg["task"]
    env = BatchedAssemblyGym(E, [H.URDF[n] for n in cfg["shapes"]], max_steps=cfg["max_steps"])
    env.reset(dict(obstacles=obstacles, targets=targets))
    dt = env.dt
    watch = (0, 1)                                       # environments whose candidate lists are compared too
    acts_log, out_log, bits_log, cand_log = [], [], [], []
    for k in range(steps):
        c = env.enumerate_actions(XG, (0.0,), amax=amax, with_bits=True)
        acts, idx = env.select_random(seed=9000 + 31 * k)
        env.sync()
        n = c["n"].cpu().numpy()
        cand = c["cand"].cpu().numpy().view(dt["action"]).reshape(E, amax)
        valid = c["valid"].cpu().numpy()
        cand_log.append({e: (cand[e][:n[e]].copy(), valid[e][:n[e]].astype(bool)) for e in watch})
        a = acts.cpu().numpy().view(dt["action"]).copy()
        i = idx.cpu().numpy()
        assert ((i >= 0) == (a["shape"] >= 0)).all()
        for e in range(E):                               # the chosen action is one of the valid candidates
            if i[e] >= 0:
                assert valid[e][i[e]] == 1 and cand[e][i[e]] == a[e]
        acts_log.append(a)
        env.step(acts)
        out_log.append(env.read_out().copy())
        bits_log.append(env.raster_bits()[0].copy())
        env.reset_done()
    assert env.candidate_overflow() == 0

    jobs = []
    for e in range(E):
        seq = [None if acts_log[k][e]["shape"] < 0 else
               tuple(int(acts_log[k][e][f]) for f in ("target_block", "target_face", "shape", "face")) +
               (float(acts_log[k][e]["offset_x"]), float(acts_log[k][e]["offset_y"])) for k in range(steps)]
        jobs.append(dict(shapes=cfg["shapes"], obstacles=obstacles, targets=targets, mu=0.8, max_steps=cfg["max_steps"],
                         actions=seq, x_ground=XG, offsets=(0.0,), cand_steps=set(range(steps)) if e in watch else set()))
    traces = H.replay_parallel(jobs)

    oenv = H.oracle_env(cfg["shapes"], obstacles, targets)
    reward_f, _ = ofeat.get_task_features(oenv.reset()[0], H.XLIM, H.YLIM, H.IMG)
    n_checked = n_band = n_implied_prev = n_resets = n_unstable = n_big = 0
    for e in range(E):
        for k in range(steps):
            ref, o, tag = traces[e][k], out_log[k][e], (case, e, k)
            if "cands" in ref:
                got_c, got_v = cand_log[k][e]
                assert len(got_c) == len(ref["cands"]), tag
                for g, w in zip(got_c, ref["cands"]):
                    assert (g["target_block"], g["target_face"], g["shape"], g["face"], g["offset_x"], g["offset_y"]) == w, tag
                assert list(got_v) == ref["cand_mask"], tag
            if ref.get("skipped"):
                assert acts_log[k][e]["shape"] < 0, tag
                continue
            assert o["error"] == 0 and o["n_blocks"] == ref["n_blocks"], tag
            assert o["n_interfaces"] == ref["n_interfaces"], tag
            assert [int(v) for v in bits_log[k][e]] == ref["bits"], tag                   # bit-exact raster
            for got, want, r_or in ((o["stable"], ref["frozen"], ref["r_frozen"]),
                                    (o["stable_unfrozen"], ref["stable_unfrozen"], ref["r_unfrozen"])):
                if r_or is not None and BAND[0] < r_or < BAND[1]:
                    n_band += 1
                    continue
                assert bool(got) == bool(want), (tag, r_or)
                n_checked += 1
                n_unstable += not want
            assert ref["stable"] == ref["frozen"], tag
            if ref["r_frozen"] is not None and BAND[0] < ref["r_frozen"] < BAND[1]:
                continue                                  # everything below follows from the frozen verdict
            assert float(o["reward"]) == ref["reward"], tag
            assert bool(o["terminated"]) == ref["terminated"] and bool(o["truncated"]) == ref["truncated"], tag
            assert o["n_targets_reached"] == ref["n_reached"], tag
            assert list(o["distance_to_targets"][:len(ref["distance"])]) == ref["distance"], tag   # bit-exact
            new_f = env.bits_to_bool(np.array(ref["new_bits"], dtype=np.uint64)).astype(np.float32)[None]
            skip_lin = ref["r_unfrozen"] is not None and BAND[0] < ref["r_unfrozen"] < BAND[1]
            if not skip_lin:
                want = float(ofeat.lin_reward(new_f, reward_f, ref["frozen"], ref["stable_unfrozen"]))
                assert abs(float(o["lin_reward"]) - want) <= 1e-5 * max(1.0, abs(want)), tag
            n_implied_prev += bool(o["solver_status"] & 4)
            n_resets += ref["terminated"] or ref["truncated"]
            n_big += ref["n_blocks"] >= 6
    # the run must have exercised what it is there for
    assert n_checked > 0.9 * 2 * E * steps * 0.9 and n_band <= 0.01 * n_checked
    assert n_implied_prev > 50 and n_resets > 2 * E and n_unstable > 100 and n_big > 20, \
        (n_implied_prev, n_resets, n_unstable, n_big)


def test_friction_change_between_steps_drops_the_carried_verdict():
    """bw_set_mu between two steps: the released-block verdict of the step before was computed with the old
    coefficient and must not stand in for the next frozen solve."""
    from bridges_b200.envs.batched import BatchedAssemblyGym
    from oracle.gym_env import Action as OAction
    # structures.py:22-30: a trapezoid lying on a slanted face, a second one on its raised slant (in equilibrium
    # iff mu > tan 60 = 1.732), then a third block far away on the floor -- its step releases the second block
    actions = [(-1, 0, 0, 0, 0.0, 0.0), (0, 3, 0, 3, 0.0, 0.0), (-1, 0, 0, 3, 3.0, 0.0)]
    env = BatchedAssemblyGym(2, [H.URDF["trapezoid"]], mu=2.0)
    env.reset(dict())
    oenv = H.oracle_env(["trapezoid"], mu=2.0)
    outs = []
    for k, a in enumerate(actions):
        if k == 2:
            env.set_mu([2.0, 0.3])                       # env 1 loses its friction before the third step
        env.step([a, a])
        outs.append(env.read_out().copy())
    assert bool(outs[1][0]["stable_unfrozen"]) and bool(outs[1][1]["stable_unfrozen"])
    for e, mu in enumerate((2.0, 0.3)):
        o = H.oracle_env(["trapezoid"], mu=2.0)
        for k, a in enumerate(actions):
            if k == 2:
                o.assembly_env.mu = mu
            obs, *_ = o.step(OAction(*a))
            frozen, unfrozen = o.stabilities_freezing()
        assert bool(outs[2][e]["stable"]) == bool(frozen), (e, outs[2][e])
        assert bool(outs[2][e]["stable_unfrozen"]) == bool(unfrozen), (e, outs[2][e])
    assert bool(outs[2][0]["stable"]) and not bool(outs[2][1]["stable"])
