"""Equilibrium verdict by non-negative least squares on the friction-cone edge rays -- TEST INFRASTRUCTURE ONLY.

The verdict statistic of this build is r* = min_{f in K} ||A f - b|| / ||b|| (the
reference asks IPOPT for feasibility of the same system, `stability.py:49-71` -> `rbe_solve`).  In 2-D every
friction cone is spanned by its two edge rays, so with R = [A(n + mu t), A(n - mu t)] (`stability.ray_matrix`)

    r* = min_{x >= 0} ||R x - b|| / ||b||,

a non-negative least-squares problem with m <= 48 rows.  This module restates the Lawson-Hanson active-set
method (Lawson & Hanson, "Solving Least Squares Problems", 1974, ch. 23) in the form a warp would run it:
the passive columns are kept as an orthogonal factorisation  Q^T R_P = [U; 0]  that grows by one Householder
reflection per added column and shrinks by Givens rotations per removed one -- no Gram matrix, no
refactorisation, O(m p) work per iteration instead of the O(m^3) Cholesky of a semismooth-Newton step.

It is a study for the next round of the CUDA solver and a second, independent check of
the oracle's residual (`stability.equilibrium_residual`, scipy BVLS) and verdict (`stability.rbe_feasible`,
HiGHS): tests/test_oracle_nnls.py.  Nothing in the product imports it.

At the solution the residual rho = b - R x satisfies R^T rho <= 0 and b . rho = ||rho||^2, so for r* > 0 it is
a Farkas vector of the system: a rigid virtual motion of the free blocks that no contact ray resists and
along which the weights do positive work.
"""
import math

import numpy as np


class NNLSResult:
    __slots__ = ("x", "residual", "resid_vec", "iterations", "removals", "chain")

    def __init__(self, x, residual, resid_vec, iterations, removals, chain):
        self.x, self.residual, self.resid_vec = x, residual, resid_vec
        self.iterations, self.removals, self.chain = iterations, removals, chain


def _givens(a, b):
    """c, s with [c s; -s c] [a; b] = [r; 0]"""
    if b == 0.0:
        return 1.0, 0.0
    r = math.hypot(a, b)
    return a / r, b / r


def nnls_rays(R, b, tol=1e-11, max_iter=None):
    """min ||R x - b|| over x >= 0 (b is used as given: normalise it outside).

    Returns NNLSResult; `iterations` counts least-squares solves on the passive set (one per added column plus
    one per removal step), `chain` the summed length of their back substitutions (the serial depth a warp sees).
    """
    R = np.asarray(R, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    m, n = R.shape
    max_iter = max_iter or 6 * n + 50
    Qt = np.eye(m)                 # Q^T, accumulated:  Qt @ R[:, order] = [U[:p, :p]; 0]
    U = np.zeros((m, m))
    qb = b.copy()                  # Q^T b
    order = []                     # passive columns in factor order
    x = np.zeros(n)
    barred = np.zeros(n, dtype=bool)      # columns found (numerically) dependent on the passive set at this point
    iterations = removals = chain = 0
    scale = max(1.0, float(np.abs(R).max())) * max(1.0, float(np.linalg.norm(b)))

    def back_substitute(p):
        s = np.zeros(p)
        for i in range(p - 1, -1, -1):
            s[i] = (qb[i] - U[i, i + 1:p] @ s[i + 1:]) / U[i, i]
        return s

    while iterations < max_iter:
        # residual of the passive-set solution without touching R: Q [0; (Q^T b)_tail]
        res = Qt[len(order):, :].T @ qb[len(order):]
        w = R.T @ res
        w[order] = -np.inf
        w[barred] = -np.inf
        j = int(np.argmax(w))
        if not (w[j] > tol * scale):
            break
        # ---- add column j: one Householder reflection on rows p.. of Q^T a_j
        p = len(order)
        if p >= m:
            break
        v = Qt @ R[:, j]
        tail = v[p:]
        norm = float(np.linalg.norm(tail))
        if norm <= 1e-12 * max(1.0, float(np.linalg.norm(v))):
            barred[j] = True           # a_j lies in the span of the passive columns: cannot enter now
            continue
        alpha = -norm if tail[0] >= 0.0 else norm
        u = tail.copy()
        u[0] -= alpha
        un = float(np.linalg.norm(u))
        if un > 0.0:
            u /= un
            Qt[p:, :] -= 2.0 * np.outer(u, u @ Qt[p:, :])
            qb[p:] -= 2.0 * u * float(u @ qb[p:])
        U[:p, p] = v[:p]
        U[p, p] = alpha
        order.append(j)
        barred[:] = False
        # ---- least squares on the passive set; walk back while a coefficient would turn non-positive
        while True:
            p = len(order)
            iterations += 1
            chain += p
            s = back_substitute(p)
            if p == 0 or float(s.min()) > 0.0:
                x[:] = 0.0
                x[order] = s
                break
            xo = x[order]
            neg = s <= 0.0
            # the newest column enters with x = 0 and a positive gradient: its s is positive, so alpha > 0
            with np.errstate(divide="ignore", invalid="ignore"):
                ratios = np.where(neg, xo / (xo - s), np.inf)
            a = float(np.min(ratios))
            xo = xo + a * (s - xo)
            drop = [k for k in range(p) if neg[k] and xo[k] <= 1e-15 * max(1.0, float(np.abs(xo).max()))]
            if not drop:
                drop = [int(np.argmin(ratios))]
            x[:] = 0.0
            x[order] = xo
            for k in sorted(drop, reverse=True):
                removals += 1
                x[order[k]] = 0.0
                # delete column k of U, restore the triangle with Givens rotations on rows (i, i + 1)
                pp = len(order)
                U[:, k:pp - 1] = U[:, k + 1:pp]
                U[:, pp - 1] = 0.0
                for i in range(k, pp - 1):
                    c, sn = _givens(U[i, i], U[i + 1, i])
                    G = np.array([[c, sn], [-sn, c]])
                    U[i:i + 2, i:pp - 1] = G @ U[i:i + 2, i:pp - 1]
                    Qt[i:i + 2, :] = G @ Qt[i:i + 2, :]
                    qb[i:i + 2] = G @ qb[i:i + 2]
                    U[i + 1, i] = 0.0
                del order[k]
    res = Qt[len(order):, :].T @ qb[len(order):]
    return NNLSResult(x, float(np.linalg.norm(qb[len(order):])), res, iterations, removals, chain)


def equilibrium_residual_nnls(A, b, mu):
    """r* of `stability.equilibrium_residual`, by Lawson-Hanson on the ray form.  Returns (r, NNLSResult)."""
    from . import stability as st
    if A.shape[0] == 0:
        return 0.0, None
    nb = float(np.linalg.norm(b))
    if nb == 0.0:
        return 0.0, None
    if A.shape[1] == 0:
        return 1.0, None
    out = nnls_rays(st.ray_matrix(A, mu), b / nb)
    return out.residual, out
