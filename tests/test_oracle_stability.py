"""Oracle self-consistency on seeded random assemblies (CPU only): HiGHS feasibility vs the
BVLS residual vs the dual Newton force solver, and the force QP against an independent
SLSQP solution."""
import numpy as np
from scipy.optimize import minimize

from oracle import stability as st
from oracle import synth


def _problems(n, seed):
    rng = np.random.default_rng(seed)
    shapes = synth.library()
    for i in range(n):
        actions = synth.random_assembly(rng, shapes, max_blocks=8)
        mu = synth.MUS[i % 3]
        env = synth.replay(actions, shapes, mu, frozen_last=bool(rng.random() < 0.5))
        asm = env.assembly_env.cra_assembly
        if asm.number_of_edges() == 0 or not asm.free_nodes():
            continue
        A, b = st.equilibrium_system(asm, mu, 1.0)
        yield A, b, mu


def test_verdict_residual_and_newton_agree():
    n = stable = 0
    for A, b, mu in _problems(120, seed=7):
        verdict = st.rbe_feasible(A, b, mu)
        r = st.equilibrium_residual(A, b, mu)
        f, y, rn, status = st.min_norm_forces(A, b, mu)
        assert -1e-7 - 1e-3 * r <= rn - r <= 2e-2 * r + 1e-7
        if 1e-9 < r < 1e-4:
            continue                      # the stated margin band
        assert verdict == (r <= 1e-9)
        assert (status == "feasible") == (r <= 1e-9)
        n += 1
        stable += bool(verdict)
    assert n > 60 and 0 < stable < n


def test_min_norm_forces_against_slsqp():
    checked = 0
    for A, b, mu in _problems(60, seed=11):
        if A.shape[1] > 24 or st.equilibrium_residual(A, b, mu) > 1e-9:
            continue
        f, _, r, status = st.min_norm_forces(A, b, mu)
        assert status == "feasible"
        ncp = A.shape[1] // 2
        cons = [dict(type="eq", fun=lambda x: A @ x - b, jac=lambda x: A)]
        for k in range(ncp):
            for sg in (1.0, -1.0):
                row = np.zeros(A.shape[1])
                row[2 * k], row[2 * k + 1] = mu, -sg
                cons.append(dict(type="ineq", fun=lambda x, row=row: row @ x, jac=lambda x, row=row: row))
        res = minimize(lambda x: 0.5 * x @ x, f + 0.01, jac=lambda x: x, constraints=cons, method="SLSQP",
                       options=dict(ftol=1e-14, maxiter=500))
        assert res.success
        assert np.linalg.norm(res.x - f) <= 1e-5 * max(1.0, np.linalg.norm(f))
        checked += 1
    assert checked >= 5
