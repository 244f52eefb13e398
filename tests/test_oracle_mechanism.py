"""The rigid-mechanism certificate (oracle/mechanism.py, CPU restatement of the CUDA `Solver::screen`)
against the LP verdict: a certificate implies HiGHS infeasibility, its vector is a Farkas vector of
the full problem, and it decides most of the infeasible systems of seeded random assemblies."""
import numpy as np

from oracle import mechanism as mech
from oracle import stability as st
from oracle import synth


def _assemblies(n, seed, max_blocks=8):
    rng = np.random.default_rng(seed)
    shapes = synth.library()
    for i in range(n):
        actions = synth.random_assembly(rng, shapes, max_blocks=max_blocks)
        mu = synth.MUS[i % 3]
        env = synth.replay(actions, shapes, mu, frozen_last=bool(rng.random() < 0.5))
        asm = env.assembly_env.cra_assembly
        if asm.number_of_edges() == 0 or not asm.free_nodes():
            continue
        yield asm, mu


def test_certificate_implies_infeasible_and_is_a_farkas_vector():
    n = infeasible = certified = 0
    for asm, mu in _assemblies(160, seed=3):
        A, b = st.equilibrium_system(asm, mu, 1.0)
        verdict = st.rbe_feasible(A, b, mu)
        cert = mech.mechanism_certificate(asm, mu)
        n += 1
        infeasible += verdict is False
        if cert is None:
            continue
        certified += 1
        assert verdict is False
        # the rigid motion of S as a dual vector y of the full system: y_j = T_j^T n for j in S
        free = asm.free_nodes()
        L0 = max([body.radius for body in asm.bodies] + [1e-300])
        y = np.zeros(A.shape[0])
        ux, uz, w = cert["motion"]
        for k, node in enumerate(free):
            if node in cert["nodes"]:
                cx, cz = asm.bodies[node + 1].com
                # virtual work of (Fx, Fz, tau_com / L0) on block j under the motion n about the origin
                y[3 * k:3 * k + 3] = (ux - w * cz / L0, uz + w * cx / L0, w)
        g = A.T @ y
        gn, gt = g[0::2], g[1::2]
        # A^T y in the dual cone of K  (gn + mu |gt| >= 0 is "no ray gains work"), b . y < 0
        assert np.all(gn - mu * np.abs(gt) >= -1e-9 * max(1.0, np.abs(g).max()))
        assert b @ y < 0.0
    assert n > 100 and infeasible > 40
    assert certified >= 0.7 * infeasible, (certified, infeasible)


def test_reference_structures():
    """utils/structures.py:22-108: the certificate never contradicts a label the LP reproduces, and it
    finds the sliding / toppling failures among them."""
    from oracle.gym_env import Action
    from tests import fixtures_structures as FS
    from tests import helpers as H
    stable = unstable = certified = 0
    for name, mu, fl, shapes, steps in FS.cases((0.8, 0.3, 2.0)):
        env = H.oracle_env(shapes, mu=mu)
        env.assembly_env.stability_fct = lambda e: (None, None)
        for a, expected in steps:
            env.step(Action(*a[:6]))
            ae = env.assembly_env
            for blk in ae.blocks:
                blk.is_static = False
            if a[6]:
                ae.blocks[-1].is_static = True
            ae._reset_cra_assembly()
            verdict = st.is_stable_rbe(ae)[0]
            cert = mech.mechanism_certificate(ae.cra_assembly, mu)
            if verdict:
                assert cert is None, (name, mu, fl)
                stable += 1
            else:
                unstable += 1
                certified += cert is not None
    assert stable > 20 and unstable > 10 and certified >= 0.6 * unstable, (stable, unstable, certified)
