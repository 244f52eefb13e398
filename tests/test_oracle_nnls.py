"""oracle/nnls.py (Lawson-Hanson on the friction-cone edge rays, incremental orthogonal factorisation) against
the oracle's two other solvers on seeded random assemblies: the residual equals scipy BVLS's, the verdict
r* <= 1e-6 equals HiGHS feasibility outside the stated band, and the residual vector of a system without
equilibrium is a Farkas vector of it."""
import numpy as np

from oracle import nnls
from oracle import stability as st
from oracle import synth

BAND = (1e-9, 1e-4)         # the excluded residual band of the GPU parity tests (tests/test_gpu_step.py)


def _systems(n, seed, max_blocks):
    rng = np.random.default_rng(seed)
    shapes = synth.library()
    for i in range(n):
        actions = synth.random_assembly(rng, shapes, max_blocks=max_blocks)
        mu = synth.MUS[i % 3]
        env = synth.replay(actions, shapes, mu, frozen_last=bool(rng.random() < 0.5))
        asm = env.assembly_env.cra_assembly
        if asm.number_of_edges() == 0 or not asm.free_nodes():
            continue
        A, b = st.equilibrium_system(asm, mu, 1.0)
        yield A, b, mu


def test_residual_verdict_and_farkas_vector():
    n = stable = unstable = in_band = 0
    iters = []
    for A, b, mu in _systems(140, seed=11, max_blocks=10):
        r, out = nnls.equilibrium_residual_nnls(A, b, mu)
        r_bvls = st.equilibrium_residual(A, b, mu)
        assert abs(r - r_bvls) <= 1e-9 + 1e-8 * r_bvls, (r, r_bvls)
        n += 1
        iters.append(out.iterations)
        R = st.ray_matrix(A, mu)
        bs = b / np.linalg.norm(b)
        y = out.resid_vec
        # KKT of min ||R x - b||, x >= 0: no ray has a positive component along the residual, x >= 0
        assert np.all(R.T @ y <= 1e-8 * max(1.0, np.abs(R).max()))
        assert np.all(out.x >= 0.0)
        if BAND[0] < r_bvls < BAND[1]:
            in_band += 1
            continue
        feasible = st.rbe_feasible(A, b, mu)
        assert (r <= 1e-6) == bool(feasible), (r, feasible)
        if feasible:
            stable += 1
        else:
            unstable += 1
            # Farkas: R^T y <= 0 (above) and b . y = ||y||^2 > 0  =>  no x >= 0 with R x = b
            assert bs @ y > 0.0 and abs(bs @ y - y @ y) <= 1e-9
    assert n > 90 and stable > 15 and unstable > 40 and in_band <= 0.02 * n
    # one least-squares solve per column that enters plus one per removal step: about one per matrix row
    assert np.mean(iters) < 30 and max(iters) < 120


def test_trivial_systems():
    A = np.zeros((3, 0))
    assert nnls.equilibrium_residual_nnls(A, np.array([0.0, 1.0, 0.0]), 0.8)[0] == 1.0
    assert nnls.equilibrium_residual_nnls(np.zeros((0, 4)), np.zeros(0), 0.8)[0] == 0.0
    # a block resting on two floor points: normal (0, 1), tangent (1, 0), weight 1 downwards -> b = (0, 1, 0)
    A = np.array([[0.0, 1.0, 0.0, 1.0], [1.0, 0.0, 1.0, 0.0], [-0.5, 0.0, 0.5, 0.0]])
    r, out = nnls.equilibrium_residual_nnls(A, np.array([0.0, 1.0, 0.0]), 0.5)
    assert r <= 1e-12 and out.iterations >= 2
