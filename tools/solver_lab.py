"""Scratch: numpy emulation of csrc/bw_solver.cuh `Solver::solve` on harvested systems, to count Newton
steps / line-search evaluations under alternative iteration strategies before spending GPU time.
Test infrastructure only (uses oracle/).

python tools/solver_lab.py /tmp/systems.pkl"""
import os, sys, pickle, math
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _project_cones(g, mu):
    """vectorised oracle.stability._project_cones; J as (ncp, 2, 2)"""
    gn, gt = g[0::2], g[1::2]
    den = 1.0 + mu * mu
    inter = np.abs(gt) <= mu * gn
    polar = (~inter) & (mu * np.abs(gt) <= -gn)
    ray = ~(inter | polar)
    sg = np.where(gt > 0, 1.0, -1.0)
    kk = (gn + mu * np.abs(gt)) / den
    f = np.zeros_like(g)
    f[0::2] = np.where(inter, gn, np.where(ray, kk, 0.0))
    f[1::2] = np.where(inter, gt, np.where(ray, sg * mu * kk, 0.0))
    J = np.zeros((gn.size, 2, 2))
    J[inter] = np.eye(2)
    u = np.stack([np.ones_like(sg), sg * mu], axis=1)
    Jr = u[:, :, None] * u[:, None, :] / den
    J[ray] = Jr[ray]
    return f, J


def _AJAt(A, J):
    m = A.shape[0]
    G = A.reshape(m, -1, 2)                       # (m, ncp, 2)
    GJ = np.einsum("mck,ckl->mcl", G, J)
    return np.einsum("mcl,ncl->mn", GJ, G)

SCHED = (1e4, 1e8, 1e8, 1e8, 1e8, 1e8)       # c_rho of csrc/bw_solver.cuh


def types_of(g, mu):
    gn, gt = g[0::2], g[1::2]
    t = np.where(np.abs(gt) <= mu * gn, 1, np.where(mu * np.abs(gt) <= -gn, 0, np.where(gt > 0, 2, 3)))
    return t


def solve(A, b, mu, sched=SCHED, r_exit=1e-6, exit_anytime=True, max_newton=60, trace=None, variant=None):
    """Mirror of Solver::solve.  Returns (status, r, newton_iters, ls_evals)."""
    m, n = A.shape
    nb = np.linalg.norm(b)
    bs = b / nb
    y = np.zeros(m)
    typ_prev = np.full(n // 2, 255)
    rprev, r = -1.0, 1.0
    status, iters, evals = 2, 0, 0
    variant = variant or {}
    for k, rho in enumerate(sched):
        inv_rho = 1.0 / rho
        yk = y.copy()
        have_r = False
        full_step = False
        for it in range(max_newton):
            g = A.T @ y
            f, J = _project_cones(g, mu)
            typ = types_of(g, mu)
            changed = bool(np.any(typ != typ_prev))
            typ_prev = typ
            rs = bs - A @ f
            grad = rs - (y - yk) * inv_rho
            gn2, rr2 = float(grad @ grad), float(rs @ rs)
            if trace is not None:
                trace.append((k, it, math.sqrt(rr2), math.sqrt(gn2), int((typ == 0).sum()), int((typ == 1).sum()),
                              int((typ >= 2).sum()), float(np.abs(y).max())))
            if gn2 <= 1e-20 or (full_step and not changed) or (exit_anytime and rr2 <= r_exit * r_exit):
                r = math.sqrt(rr2)
                have_r = True
                break
            if variant.get("farkas"):
                # candidate certificate z = rs: A^T z in the polar cone up to eps, b.z > delta |z|
                z = rs
                w = A.T @ z
                pw, _ = _project_cones(w, mu)
                nz = math.sqrt(rr2)
                if np.abs(pw).sum() <= variant["farkas"] * nz and float(bs @ z) > 1e-3 * nz:
                    return 1, math.sqrt(rr2), iters, evals
            H = _AJAt(A, J)
            H[np.diag_indices(m)] += inv_rho
            d = np.linalg.solve(H, grad)
            h = A.T @ d
            gd, dd, fh0 = float(grad @ d), float(d @ d), float(f @ h)
            phi0 = gd
            if not (phi0 > 1e-30):
                break
            base = phi0 + fh0

            def fdoth(t):
                ft, _ = _project_cones(g + t * h, mu)
                return float(ft @ h)
            t = 1.0
            p = base - fdoth(1.0) - dd * inv_rho
            evals += 1
            if p < -variant.get('accept', 0.1) * phi0:   # the kernel keeps a full step that passes the search's own test
                lo, plo, hi, phi = 0.0, phi0, 1.0, p
                for _ls in range(20):
                    w = hi - lo
                    t = lo + w * plo / (plo - phi)
                    t = min(max(t, lo + 0.1 * w), hi - 0.1 * w)
                    p = base - fdoth(t) - t * dd * inv_rho
                    evals += 1
                    if abs(p) <= 0.1 * phi0:
                        break
                    if p > 0.0:
                        lo, plo = t, p
                    else:
                        hi, phi = t, p
                if p < 0.0 and abs(p) > 0.1 * phi0 and lo > 0.0:
                    t = lo
            y = y + t * d
            iters += 1
            full_step = (t == 1.0)
            ymax = float(np.abs(y).max())
            if t * t * dd <= 1e-30 * max(1.0, ymax * ymax):
                break
        if not have_r:
            g = A.T @ y
            f, _ = _project_cones(g, mu)
            typ_prev = types_of(g, mu)
            r = float(np.linalg.norm(bs - A @ f))
        if r <= r_exit:
            status = 0
            break
        if rprev >= 0.0 and abs(r - rprev) <= 1e-3 * r:
            status = 1
            break
        if rprev >= 0.0 and r >= 0.9 * rprev and r > 1e-3:
            status = 1
            break
        rprev = r
    return status, r, iters, evals


def load(path):
    with open(path, "rb") as fh:
        recs = pickle.load(fh)
    systems = []
    for r in recs:
        for tag in ("frozen", "unfrozen"):
            if r[tag] is None:
                continue
            A, b, ok = r[tag]
            if A.shape[0] == 0 or A.shape[1] == 0 or not np.any(b):
                continue
            systems.append((A, b, r["mu"], ok, r["n_blocks"], tag))
    return systems


def evaluate(systems, name, **kw):
    rows = []
    bad = 0
    for (A, b, mu, ok, nbk, tag) in systems:
        st, r, it, ev = solve(A, b, mu, **kw)
        verdict = (st == 0) or (st == 2 and r <= 1e-6)
        if ok is not None and verdict != ok:
            bad += 1
        rows.append((nbk, ok, st, it, ev))
    rows = np.array([(a, -1 if b is None else int(b), c, d, e) for a, b, c, d, e in rows])
    it = rows[:, 3]
    print(f"{name:34s} n={len(rows)} mismatches={bad} iters mean {it.mean():.2f} p99 {np.percentile(it, 99):.0f} max {it.max()}"
          f" | stable mean {it[rows[:, 1] == 1].mean():.2f} max {it[rows[:, 1] == 1].max()}"
          f" | unstable mean {it[rows[:, 1] == 0].mean():.2f} max {it[rows[:, 1] == 0].max()}"
          f" | ls evals mean {rows[:, 4].mean():.2f}")
    big = rows[rows[:, 0] >= 7]
    if len(big):
        print(f"{'':34s} >=7 blocks n={len(big)}: iters mean {big[:, 3].mean():.2f} max {big[:, 3].max()}"
              f" | unstable mean {big[big[:, 1] == 0][:, 3].mean():.2f}")
    return rows


if __name__ == "__main__":
    systems = load(sys.argv[1])
    print(len(systems), "systems")
    evaluate(systems, "baseline (CUDA schedule)")
