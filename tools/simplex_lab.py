"""Scratch: numpy emulation of a warm-started phase-1 simplex verdict solver on harvested episodes
(tools/harvest_episodes.py): how many pivots does a step need when the basis of the previous step is kept?
Test infrastructure only (reads oracle-made data).  The tidy form of the algorithm, with the certificates and the
torque-scale similarity, is oracle/simplex.py (tests/test_oracle_simplex.py); this file keeps the pivot-rule variants
that were compared.

Problem of a step: rays r_{c,+-} = a_n(c) +- mu a_t(c) of every contact point c, find lambda >= 0 with R lambda = b
(= A f = b, f in K).  Phase 1: min sum of artificials.  Basis identity across steps: ray = (a, b, point, sign)
with (a, b) the interface's bodies; artificial = (block, k, sign).

python tools/simplex_lab.py EPISODES.pkl"""
import sys, pickle, math, collections
import numpy as np

PIV_TOL = 1e-7       # smallest pivot element accepted in the ratio test
D_TOL = 1e-9         # reduced-cost tolerance
Z_TOL = 1e-7         # phase-1 objective under which the system counts as feasible
MAX_PIV = 400
HARRIS = 1e-9        # feasibility tolerance of the Harris ratio test


class Basis:
    """ids: list of basic column ids; Binv; xB"""
    def __init__(self):
        self.ids, self.Binv, self.xB = [], np.zeros((0, 0)), np.zeros(0)
        self.rows = []            # row ids (block, k) in row order


def ray_columns(A, itf, free, mu):
    """dict id -> column; ids (a, b, q, s)"""
    cols = {}
    for i, (a, b) in enumerate(itf):
        for q in range(2):
            c = 2 * i + q
            an, at = A[:, 2 * c], A[:, 2 * c + 1]
            if not (np.any(an) or np.any(at)):
                continue
            cols[(a, b, q, +1)] = an + mu * at
            cols[(a, b, q, -1)] = an - mu * at
    return cols


def run_phase1(cols, rows, bvec, basis, stats, rule="dantzig"):
    """continue phase 1 from `basis` (whose rows must equal `rows`).  Returns (feasible, pivots, z, ok)"""
    m = len(rows)
    ids, Binv, xB = basis.ids, basis.Binv, basis.xB
    keys = list(cols.keys())
    Rm = np.array([cols[k] for k in keys]).T if keys else np.zeros((m, 0))
    in_basis = {k: i for i, k in enumerate(ids)}
    piv = 0
    while True:
        cB = np.array([1.0 if k[0] == "art" else 0.0 for k in ids])
        z = float(cB @ xB)
        if z <= Z_TOL:
            return True, piv, z, True
        pi = cB @ Binv
        d = -(pi @ Rm)
        for k, i in in_basis.items():
            if k[0] != "art":
                d[keys.index(k)] = 0.0
        if rule == "gpu":
            # most negative reduced cost through a float key with its low 8 bits replaced by the ray index
            df = d.astype(np.float32)
            u = df.view(np.uint32).astype(np.uint64)
            key = np.where(u & 0x80000000, (~u) & 0xffffffff, u | 0x80000000)
            key = (key & 0xffffff00) | np.arange(len(d), dtype=np.uint64)
            key = np.where(d < 0, key, 0xffffffff)
            q = int(np.argmin(key))
        elif rule == "dantzig":
            q = int(np.argmin(d))
        elif rule == "bland":
            neg = np.nonzero(d < -D_TOL)[0]
            q = int(neg[0]) if len(neg) else int(np.argmin(d))
        else:       # steepest-ish: scale by column norm in the current basis is too dear; use |r_j|
            q = int(np.argmin(d / np.maximum(np.linalg.norm(Rm, axis=0), 1e-300)))
        if d[q] >= -D_TOL:
            return False, piv, z, True
        w = Binv @ Rm[:, q]
        cand = w > PIV_TOL
        if not np.any(cand):
            return False, piv, z, False          # numerically unbounded: failure
        # Harris ratio test: the bound from the relaxed ratios (x_i + delta) / w_i, then the largest pivot
        # element among the rows whose plain ratio is under that bound
        wc = np.where(cand, w, 1.0)
        tmax = np.where(cand, (xB + HARRIS) / wc, np.inf).min()
        theta = np.where(cand, xB / wc, np.inf)
        near = cand & (theta <= tmax)
        p = int(np.argmax(np.where(near, w, -np.inf)))
        if rule == "gpu":
            rel = np.where(cand, (xB + HARRIS) / wc, np.inf)
            tm = float(np.nextafter(np.float32(rel.min()), np.float32(np.inf)))      # >= the float rounded up
            ok = cand & (xB <= tm * w)
            kw = (w.astype(np.float32).view(np.uint32).astype(np.uint64) & 0xffffffc0) | np.arange(len(w), dtype=np.uint64)
            p = int(np.argmax(np.where(ok, kw, 0)))
        # update
        wp = w[p]
        rowp = Binv[p] / wp
        Binv = Binv - np.outer(w, rowp)
        Binv[p] = rowp
        xp = xB[p] / wp
        xB = xB - w * xp
        xB[p] = xp
        xB = np.maximum(xB, 0.0)
        del in_basis[ids[p]]
        ids[p] = keys[q]
        in_basis[keys[q]] = p
        basis.Binv, basis.xB = Binv, xB
        piv += 1
        if piv >= MAX_PIV:
            return False, piv, z, False


def extend_rows(basis, new_rows, bnew):
    """add rows with artificial basics (old basic columns have no entries in the new rows)"""
    m0, k = len(basis.rows), len(new_rows)
    Binv = np.zeros((m0 + k, m0 + k))
    Binv[:m0, :m0] = basis.Binv
    xB = np.concatenate([basis.xB, np.abs(bnew)])
    for i, r in enumerate(new_rows):
        s = 1.0 if bnew[i] >= 0 else -1.0
        Binv[m0 + i, m0 + i] = s
        basis.ids.append(("art", r, s))
    basis.Binv, basis.xB = Binv, xB
    basis.rows = basis.rows + list(new_rows)


def copy_basis(b):
    c = Basis()
    c.ids, c.Binv, c.xB, c.rows = list(b.ids), b.Binv.copy(), b.xB.copy(), list(b.rows)
    return c


def main(path, rule="dantzig", limit=None):
    recs = pickle.load(open(path, "rb"))
    if limit:
        recs = recs[:limit]
    print(len(recs), "steps")
    stats = collections.defaultdict(list)
    mism = 0
    R = None            # basis of the released problem of the previous step
    R_feasible = False
    prev_ep = None
    for r in recs:
        n = r["n_blocks"]
        if r["episode"] != prev_ep or R is None:
            R = Basis()
            R.scale = 1.0
            R_feasible = True
            prev_ep = r["episode"]
        if r["A"] is None:
            R = None
            continue
        A, b, mu = r["A"], r["b"], r["mu"]
        free = r["free"]
        assert free == list(range(n)), free
        nb = np.linalg.norm(b)
        bs = b / nb
        rows_all = [(blk, k) for blk in free for k in range(3)]
        # rows the stored basis knows about
        known = len(R.rows)
        if known > 3 * (n - 1):
            R = Basis(); known = 0; R_feasible = True
        # the stored x_B is in units of the previous normalisation: rescale
        if known:
            R.xB = R.xB * (R.scale / nb)
        # ---- frozen(t): rows of blocks 0..n-2
        mF = 3 * (n - 1)
        if mF > 0:
            colsF = {k: v[:mF] for k, v in ray_columns(A, r["itf"], free, mu).items() if np.any(v[:mF])}
            if known == mF and R_feasible:
                stats["frozen_implied"].append(0)
                okF = True
            else:
                Fb = copy_basis(R)
                if known < mF:
                    extend_rows(Fb, rows_all[known:mF], bs[known:mF])
                okF, piv, z, fine = run_phase1(colsF, rows_all[:mF], bs[:mF], Fb, stats, rule)
                stats["frozen_piv"].append(piv)
                stats["frozen_piv_nb"].append((n, piv, int(okF)))
                if not fine:
                    stats["fail"].append(1)
            if r["frozen_ok"] is not None and bool(r["frozen_ok"]) != okF:
                mism += 1
                stats["mism"].append(("F", n, r["frozen_ok"], okF))
        # ---- released(t): all rows
        m = 3 * n
        cols = ray_columns(A, r["itf"], free, mu)
        Rb = R
        if known < m:
            extend_rows(Rb, rows_all[known:m], bs[known:m])
        okR, piv, z, fine = run_phase1(cols, rows_all, bs, Rb, stats, rule)
        stats["released_piv"].append(piv)
        stats["released_piv_nb"].append((n, piv, int(okR)))
        if not fine:
            stats["fail"].append(1)
        if r["released_ok"] is not None and bool(r["released_ok"]) != okR:
            mism += 1
            stats["mism"].append(("R", n, r["released_ok"], okR))
        R_feasible = okR
        R.scale = nb
    for key in ("frozen_piv", "released_piv"):
        v = np.array(stats[key])
        print(f"{key:14s} n={len(v)} mean {v.mean():.2f} p50 {np.percentile(v, 50):.0f} p90 {np.percentile(v, 90):.0f} "
              f"p99 {np.percentile(v, 99):.0f} max {v.max()}")
    print("frozen implied", len(stats["frozen_implied"]), "failures", len(stats["fail"]), "verdict mismatches", mism)
    print(stats["mism"][:20])
    for key in ("frozen_piv_nb", "released_piv_nb"):
        v = np.array(stats[key])
        for nbk in sorted(set(v[:, 0])):
            s = v[v[:, 0] == nbk]
            print(f"  {key} n_blocks={nbk:2d} n={len(s):5d} piv mean {s[:, 1].mean():5.2f} p99 {np.percentile(s[:, 1], 99):4.0f} max {s[:, 1].max():3d}"
                  f" feasible {s[:, 2].mean():.2f}")


if __name__ == "__main__":
    main(sys.argv[1], *(sys.argv[2:3]))
