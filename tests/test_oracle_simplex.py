"""oracle/simplex.py -- the CPU restatement of the CUDA LP verdict path (csrc/bw_lp.cuh: warm-started phase-1 simplex
with the kernel's selection rules, tolerances and certificates) -- against the HiGHS verdicts of the oracle
(`oracle.stability.rbe_feasible`, the restated `is_stable_rbe`, stability.py:49-71) along whole episodes of the bench
rollouts: every certified verdict is the oracle's outside the residual band, almost every run is certified, and the
warm start keeps the pivot counts where the kernel's cost model has them."""
import numpy as np
import pytest

from oracle import actions as oact
from oracle import features as ofeat
from oracle import simplex as sx
from oracle import stability as ost
from oracle.rendering import render_blocks_2d
from tests import helpers as H

BAND = (1e-9, 1e-4)
XG = [-2.0 + 2.0 * i / 9 for i in range(10)]
XLIM, YLIM, IMG = (-3.0, 7.0), (0.0, 10.0), (64, 64)


def _episodes(shapes, obstacles, targets, max_steps, n_steps, seed):
    """random valid rollouts on the oracle env (the bench's policy); yields per step the system with every block
    released, the interface list and the oracle's two verdicts with their residuals"""
    rng = np.random.default_rng(seed)
    env = H.oracle_env(shapes, obstacles, targets, mu=0.8, max_steps=max_steps)
    done_steps = 0
    while done_steps < n_steps:
        obs, _ = env.reset()
        obstacle_f = render_blocks_2d(obs['obstacle_blocks'], XLIM, YLIM, IMG).astype(np.float32)[None]
        first = True
        done = False
        while not done and done_steps < n_steps:
            block_f, _ = ofeat.get_state_features(obs, XLIM, YLIM, IMG)
            cands = [*oact.generate_actions(env, XG, [0.0])]
            cand_f = ofeat.get_action_features(env, cands, XLIM, YLIM, IMG)
            kept, _, _ = oact.filter_actions(env, cands, cand_f, block_f, obstacle_f, XLIM, YLIM)
            if not kept:
                break
            obs, _, terminated, truncated, _ = env.step(kept[int(rng.integers(len(kept)))])
            ae = env.assembly_env
            n = len(ae.blocks)
            frozen, released = env.stabilities_freezing()
            r_frozen, r_released = H.residuals(env)
            ae.unfreeze_block(n - 1)
            asm = ae.cra_assembly
            A, b = ost.equilibrium_system(asm, ae.mu, ae.density)
            itf = [(it.a, it.b) for it in asm.interfaces]
            L0 = max([body.radius for body in asm.bodies] + [1e-300])
            ae.freeze_block(n - 1)
            yield dict(new_episode=first, n=n, A=A, b=b, itf=itf, mu=ae.mu, L0=L0, frozen=frozen, released=released,
                       r_frozen=r_frozen, r_released=r_released)
            first = False
            done_steps += 1
            done = bool(terminated or truncated)


@pytest.mark.parametrize("case", ["tower2", "bridge5_mixed"])
def test_warm_started_simplex_reaches_the_oracle_verdicts(case):
    sq = 0.6
    if case == "tower2":
        shapes, obstacles, targets, max_steps = ["trapezoid"], [(sq, 0, sq / 2)], [(sq, 0, sq + sq / 2)], 10
    else:
        shapes, max_steps = ["trapezoid", "hexagon"], 15
        obstacles, targets = [(i * sq, 0, sq / 2) for i in range(1, 6)], [(5 * sq + 2.5 * sq, 0, sq / 2)]
    ev = None
    runs = certified = in_band = 0
    pivots = []
    for rec in _episodes(shapes, obstacles, targets, max_steps, n_steps=220, seed=11):
        if rec["new_episode"]:
            ev = sx.EpisodeVerdicts()
        out = ev.step(rec["A"], rec["b"], rec["itf"], rec["mu"], rec["n"], L0=rec["L0"])
        for tag, want, r_or in (("frozen", rec["frozen"], rec["r_frozen"]), ("released", rec["released"], rec["r_released"])):
            verdict, piv, resid, implied = out[tag]
            if rec["n"] == 1 and tag == "frozen":
                continue                                         # nothing is free
            if not implied:
                runs += 1
                pivots.append(piv)
                certified += verdict != sx.NOT_CERTIFIED
            if verdict == sx.NOT_CERTIFIED:
                continue
            if r_or is not None and BAND[0] < r_or < BAND[1]:
                in_band += 1
                continue
            assert (verdict == sx.FEASIBLE) == bool(want), (case, tag, rec["n"], piv, resid, r_or)
            if verdict == sx.FEASIBLE and resid is not None:
                assert resid <= 1e-6                             # the certificate: a basic solution inside the cones
    pivots = np.array(pivots)
    assert runs > 250 and certified >= runs - 2, (runs, certified)
    assert in_band <= 0.02 * runs
    assert pivots.mean() < 5.0 and pivots.max() <= 40, (pivots.mean(), pivots.max())


def test_empty_basis_and_cap():
    """a problem without history starts from the all-artificial basis: same verdicts, about 1.5 m pivots"""
    from oracle import synth
    rng = np.random.default_rng(5)
    lib = synth.library()
    n_checked = 0
    for i in range(40):
        actions = synth.random_assembly(rng, lib, max_blocks=8, min_blocks=3)
        mu = synth.MUS[i % 3]
        env = synth.replay(actions, lib, mu, frozen_last=False)
        asm = env.assembly_env.cra_assembly
        if asm.number_of_edges() == 0 or not asm.free_nodes():
            continue
        A, b = ost.equilibrium_system(asm, mu, 1.0)
        want = ost.rbe_feasible(A, b, mu)
        r = ost.equilibrium_residual(A, b, mu)
        cols = sx.ray_columns(A, [(it.a, it.b) for it in asm.interfaces], mu)
        basis = sx.Basis()
        bs = b / np.linalg.norm(b)
        sx.extend_rows(basis, len(b), bs)
        verdict, piv, resid = sx.run_phase1(cols, bs, basis)
        assert piv <= 2 * len(b) + 24
        if verdict == sx.NOT_CERTIFIED or BAND[0] < r < BAND[1] or want is None:
            continue
        assert (verdict == sx.FEASIBLE) == bool(want), (i, piv, r)
        n_checked += 1
    assert n_checked >= 25
