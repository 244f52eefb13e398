"""CPU restatement of the warm-started LP verdict path of the CUDA step kernel (csrc/bw_lp.cuh) -- TEST
INFRASTRUCTURE ONLY (see oracle/__init__.py).

The question is the one `is_stable_rbe` asks (assembly_gym/utils/stability.py:49-71): is there f in the friction
cones with A f = b.  In the 2-D reduction the cones are polyhedral, f = lambda+ (1, +mu) + lambda- (1, -mu) per contact
point, so it is the feasibility of  R lambda = b, lambda >= 0  with the cone edge rays as columns
(`oracle.stability.ray_matrix`).  This module follows the kernel step by step:

  * phase 1 of the revised simplex method with an explicit basis inverse, artificial columns +e_i (b >= 0: weights);
  * Dantzig pricing through a float key that carries the ray index (ties -> lowest ray), Harris ratio test with the
    bound rounded up to float and the largest pivot element under it (ties -> highest row), the kernel's tolerances;
  * the warm start along an episode: the final basis of the released problem of step t-1 (all blocks free) is the
    start of both problems of step t -- old rows keep their basic columns, the rows of new blocks start with their
    artificial column; columns are identified by (body a, body b, contact point, ray sign);
  * every verdict certified against the system itself: feasible = a basic solution with ||b - R lambda|| <= tol,
    infeasible = the dual vector pi of the optimal basis with  pi . r <= CERT_REL pi . b  for EVERY ray and pi . b
    above Z_INF; anything else is "not certified" (the kernel then runs its Newton solver).

`tests/test_oracle_simplex.py` checks it against the HiGHS verdict of `oracle.stability.rbe_feasible` on whole oracle
episodes; `tools/simplex_lab.py` uses it for pivot statistics on harvested episodes."""
import numpy as np

PIV_TOL = 1e-7       # smallest pivot element
HARRIS = 1e-9        # feasibility slack of the ratio test
D_TOL = 1e-9         # reduced costs above -D_TOL count as non-negative
CERT_REL = 1e-5      # certificate of infeasibility: pi . r <= CERT_REL pi . b for every ray
FEASIBLE, INFEASIBLE, NOT_CERTIFIED = 1, 2, 0


class Basis:
    """ids: basic column of every row position -- ("art", row) or (a, b, point, sign); Binv; xB; rows covered"""
    def __init__(self):
        self.ids, self.Binv, self.xB, self.m = [], np.zeros((0, 0)), np.zeros(0), 0

    def copy(self):
        c = Basis()
        c.ids, c.Binv, c.xB, c.m = list(self.ids), self.Binv.copy(), self.xB.copy(), self.m
        return c


def ray_columns(A, interfaces, mu):
    """ordered dict ray id -> column.  interfaces: (a, b) body pairs in the order of A's column groups"""
    cols = {}
    for i, (a, b) in enumerate(interfaces):
        for q in range(2):
            c = 2 * i + q
            an, at = A[:, 2 * c], A[:, 2 * c + 1]
            cols[(a, b, q, +1)] = an + mu * at
            cols[(a, b, q, -1)] = an - mu * at
    return cols


def extend_rows(basis, m, b):
    """rows basis.m .. m-1 join with their artificial columns (old basic columns have no entries there)"""
    m0 = basis.m
    if m <= m0:
        return
    Binv = np.zeros((m, m))
    Binv[:m0, :m0] = basis.Binv
    Binv[m0:, m0:] = np.eye(m - m0)
    basis.Binv = Binv
    basis.xB = np.concatenate([basis.xB, b[m0:m]])
    basis.ids = basis.ids + [("art", i) for i in range(m0, m)]
    basis.m = m


def _ordered_key(x):
    """monotone float32 -> uint32 map of the kernel (`ordered_key`)"""
    u = np.asarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    return np.where(u & 0x80000000, (~u) & 0xffffffff, u | 0x80000000)


def run_phase1(cols, b, basis, r_exit=1e-6, z_inf=1e-5):
    """Phase 1 from `basis` (rows = len(b)).  Returns (verdict, pivots, residual or None)."""
    m = len(b)
    keys = list(cols.keys())
    R = np.array([cols[k][:m] for k in keys]).T if keys else np.zeros((m, 0))
    ids, Binv, xB = basis.ids, basis.Binv, basis.xB
    pos = {k: i for i, k in enumerate(ids)}
    pivots, maxpiv = 0, 2 * m + 24
    blocked = False

    def art():
        return np.array([1.0 if k[0] == "art" else 0.0 for k in ids])

    while True:
        cB = art()
        z = float(cB @ xB)
        if z <= r_exit:
            lam = np.zeros(len(keys))
            for j, k in enumerate(keys):
                if k in pos:
                    lam[j] = max(xB[pos[k]], 0.0)
            r = float(np.linalg.norm(b - R @ lam))
            basis.Binv, basis.xB = Binv, xB
            return (FEASIBLE, pivots, r) if r <= r_exit else (NOT_CERTIFIED, pivots, r)
        if pivots >= maxpiv:
            return NOT_CERTIFIED, pivots, None
        pi = cB @ Binv
        d = -(pi @ R)
        nonbasic = np.array([k not in pos for k in keys]) if keys else np.zeros(0, dtype=bool)
        cand = nonbasic & (d < 0.0)
        q = -1
        if np.any(cand):
            key = (_ordered_key(d) & 0xffffff00) | (np.arange(len(d), dtype=np.uint64) & 0xff)
            key = np.where(cand, key, 0xffffffff)
            q = int(np.argmin(key))
        dq = d[q] if q >= 0 else 0.0
        optimal = not (dq < -D_TOL) or blocked
        # the certificate holds for any vector pi, so it is also tried as soon as the most negative reduced cost is
        # small against the objective (what is left to gain is then rounding noise, not a direction of descent)
        if optimal or (z > z_inf and dq >= -CERT_REL * z):
            pb = float(pi @ b)
            dmin = float(d.min()) if len(d) else 0.0
            if pb > z_inf and min(dmin, 0.0) >= -CERT_REL * pb:
                basis.Binv, basis.xB = Binv, xB
                return INFEASIBLE, pivots, None
            if optimal:
                basis.Binv, basis.xB = Binv, xB
                return NOT_CERTIFIED, pivots, None
        w = Binv @ R[:, q]
        ok = w > PIV_TOL
        if not np.any(ok):
            if dq < -1e-6:
                return NOT_CERTIFIED, pivots, None
            blocked = True
            continue
        wc = np.where(ok, w, 1.0)
        rel = np.where(ok, (xB + HARRIS) / wc * (1.0 + 1e-9), np.inf)
        tmax = float(np.nextafter(np.float32(rel.min()), np.float32(np.inf)))     # the bound rounded up to float
        under = ok & (xB <= tmax * w)
        kw = (w.astype(np.float32).view(np.uint32).astype(np.uint64) & 0xffffffc0) | np.arange(m, dtype=np.uint64)
        p = int(np.argmax(np.where(under, kw, 0)))
        wp = w[p]
        rowp = Binv[p] / wp
        theta = xB[p] / wp
        Binv = Binv - np.outer(np.where(np.arange(m) == p, 0.0, w), rowp)
        Binv[p] = rowp
        xB = np.maximum(xB - w * theta, 0.0)
        xB[p] = theta
        del pos[ids[p]]
        ids[p] = keys[q]
        pos[keys[q]] = p
        pivots += 1


class EpisodeVerdicts:
    """The step logic of the kernel's LP phase along one episode (csrc/bw_step.cu, phase 3a): call `step` after
    every placement with the equilibrium system of ALL blocks released (the frozen problem is that system without
    the last block's three rows)."""
    def __init__(self):
        self.R = Basis()             # final basis of the released problem of the last step
        self.R_feasible = True       # (an empty assembly is in equilibrium)
        self.scale = 1.0
        self.L0 = None               # torque-row scale the stored inverse was built with

    def _rescale(self, L0):
        """A block with a larger radius changes the torque-row scale L0 of `equilibrium_system` (mixed libraries): the
        torque rows of every ray column are s = L0_old / L0 times what the stored inverse was built with, B' = D B E
        with D = diag(1, 1, s, ...) and E = 1/s on the artificial columns of torque rows (unit vectors in either
        scaling), so B'^-1 = E^-1 B^-1 D^-1 and, the torque rows of b being zero, x' = E^-1 x  (`Lp::setup`)."""
        R = self.R
        if R.m == 0 or self.L0 is None or self.L0 == L0:
            return
        s = self.L0 / L0
        torque = np.arange(R.m) % 3 == 2
        R.Binv[:, torque] /= s
        for i, k in enumerate(R.ids):
            if k[0] == "art" and i % 3 == 2:
                R.Binv[i, :] *= s
                R.xB[i] *= s

    def step(self, A, b, interfaces, mu, n_blocks, L0=None):
        """-> dict(frozen=(verdict, pivots, residual, implied), released=(...)); verdicts FEASIBLE / INFEASIBLE /
        NOT_CERTIFIED.  L0: the torque-row scale of A (largest body radius), needed when it changes along the episode"""
        if L0 is not None:
            self._rescale(L0)
            self.L0 = L0
        nb = float(np.linalg.norm(b))
        bs = b / nb
        m, mF = 3 * n_blocks, 3 * (n_blocks - 1)
        cols = ray_columns(A, interfaces, mu)
        if self.R.m > mF:                                   # (cannot happen along real steps)
            self.R, self.R_feasible = Basis(), True
        if self.R.m:
            # the stored basic solution in the normalisation of this step
            self.R.xB = np.maximum(self.R.xB * (self.scale / nb), 0.0)
        out = {}
        if mF == 0:
            out["frozen"] = (FEASIBLE, 0, 0.0, True)        # nothing is free
        elif self.R.m == mF and self.R_feasible:
            out["frozen"] = (FEASIBLE, 0, None, True)       # implied by the previous released verdict
        else:
            F = self.R.copy()
            # the frozen problem is normalised by the weights of ITS free blocks
            nF = float(np.linalg.norm(b[:mF]))
            F.xB = F.xB * (nb / nF)
            bF = b[:mF] / nF
            extend_rows(F, mF, bF)
            v, piv, r = run_phase1({k: c[:mF] for k, c in cols.items() if np.any(c[:mF])}, bF, F)
            out["frozen"] = (v, piv, r, False)
        if out["frozen"][0] == INFEASIBLE:
            out["released"] = (INFEASIBLE, 0, None, True)   # no frozen equilibrium => no released one; episode over
            self.R, self.R_feasible = Basis(), True
            return out
        if out["frozen"][0] == NOT_CERTIFIED:
            out["released"] = (NOT_CERTIFIED, 0, None, False)
            self.R, self.R_feasible = Basis(), True
            return out
        extend_rows(self.R, m, bs)
        v, piv, r = run_phase1(cols, bs, self.R)
        out["released"] = (v, piv, r, False)
        if v == NOT_CERTIFIED:
            self.R, self.R_feasible = Basis(), True
        else:
            self.R_feasible = v == FEASIBLE
            self.scale = nb
        return out
