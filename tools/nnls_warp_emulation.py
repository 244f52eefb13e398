"""Scratch: lane-by-lane Python transcription of tools/ubench/nnls_warp.cu (same arrays, same order of operations,
same selections) to check the kernel's logic on the CPU against oracle/nnls.py before it first runs on a GPU.
Test infrastructure only.

python tools/nnls_warp_emulation.py tools/ubench/systems_400.bin"""
import math, struct, sys
import numpy as np

MR, MC = 32, 128


def emulate(m, n, b, cols):
    """cols[j] = list of (row, value).  Returns (residual, iterations)."""
    Qt = np.zeros((MR, MR)); U = np.zeros((MR, MR))
    for i in range(MR):
        Qt[i, i] = 1.0
    qb = np.zeros(MR); qb[:m] = b
    xo = np.zeros(MR); invd = np.zeros(MR)
    order = [0] * MR
    inP = [False] * MC; barred = [False] * MC
    scale = max(1.0, max(abs(v) for c in cols for _, v in c))
    tol = 1e-11 * scale
    p = iters = 0
    max_iter = 6 * n + 50
    while iters < max_iter and p < m:
        res = np.zeros(MR)
        for lane in range(m):
            res[lane] = sum(Qt[i, lane] * qb[i] for i in range(p, m))
        best, bj = -math.inf, -1
        for j in range(n):                       # lowest index among equal maxima, like the butterfly's tie rule
            if inP[j] or barred[j]:
                continue
            w = sum(v * res[r] for r, v in cols[j])
            if w > best:
                best, bj = w, j
        if not (best > tol) or bj < 0:
            break
        v = np.zeros(MR)
        for lane in range(m):
            v[lane] = sum(val * Qt[lane, r] for r, val in cols[bj])
        vv = float(v @ v)
        tail2 = float(v[p:m] @ v[p:m])
        norm = math.sqrt(tail2)
        if norm <= 1e-12 * max(1.0, math.sqrt(vv)):
            barred[bj] = True
            continue
        alpha = -norm if v[p] >= 0.0 else norm
        u = np.zeros(MR)
        u[p:m] = v[p:m]
        u[p] -= alpha
        un2 = float(u @ u)
        if un2 > 0.0:
            u = u / math.sqrt(un2)
            for lane in range(m):                # lane = column of Q^T
                t = 2.0 * sum(u[i] * Qt[i, lane] for i in range(p, m))
                for i in range(p, m):
                    Qt[i, lane] -= u[i] * t
            dot = 2.0 * float(u[:m] @ qb[:m])
            for lane in range(p, m):
                qb[lane] -= u[lane] * dot
        for lane in range(p):
            U[lane, p] = v[lane]
        U[p, p] = alpha; invd[p] = 1.0 / alpha; order[p] = bj; inP[bj] = True; xo[p] = 0.0
        barred = [False] * MC
        p += 1
        while True:
            iters += 1
            acc = np.zeros(MR); acc[:p] = qb[:p]
            s = np.zeros(MR)
            for k in range(p - 1, -1, -1):
                sk = acc[k] * invd[k]
                s[k] = sk
                for lane in range(k):
                    acc[lane] -= U[lane, k] * sk
            if p == 0 or s[:p].min() > 0.0:
                xo[:p] = s[:p]
                break
            x = xo.copy()
            neg = [(lane < p) and s[lane] <= 0.0 for lane in range(MR)]
            ratio = [x[l] / (x[l] - s[l]) if neg[l] else math.inf for l in range(MR)]
            a = min(ratio)
            for lane in range(p):
                x[lane] += a * (s[lane] - x[lane])
            xmax = max([abs(x[l]) for l in range(p)] + [0.0])
            drop = [l for l in range(MR) if neg[l] and x[l] <= 1e-15 * max(1.0, xmax)]
            if not drop:
                drop = [min(l for l in range(MR) if neg[l] and ratio[l] == a)]
            xo[:p] = x[:p]
            for k in sorted(drop, reverse=True):
                col = order[k]
                for lane in range(m):
                    for c in range(k, p - 1):
                        U[lane, c] = U[lane, c + 1]
                    U[lane, p - 1] = 0.0
                xn = [xo[l + 1] if l + 1 < p else 0.0 for l in range(MR)]
                on = [order[l + 1] if l + 1 < p else 0 for l in range(MR)]
                for lane in range(k, p):
                    xo[lane] = xn[lane]; order[lane] = on[lane]
                inP[col] = False
                for i in range(k, p - 1):
                    ga, gb = U[i, i], U[i + 1, i]
                    c, sn = 1.0, 0.0
                    if gb != 0.0:
                        rr = math.sqrt(ga * ga + gb * gb); c = ga / rr; sn = gb / rr
                    for lane in range(i, p - 1):
                        a0, a1 = U[i, lane], U[i + 1, lane]
                        U[i, lane] = c * a0 + sn * a1
                        U[i + 1, lane] = 0.0 if lane == i else c * a1 - sn * a0
                    for lane in range(m):
                        q0, q1 = Qt[i, lane], Qt[i + 1, lane]
                        Qt[i, lane] = c * q0 + sn * q1
                        Qt[i + 1, lane] = c * q1 - sn * q0
                    b0, b1 = qb[i], qb[i + 1]
                    qb[i] = c * b0 + sn * b1
                    qb[i + 1] = c * b1 - sn * b0
                    invd[i] = 1.0 / U[i, i]
                p -= 1
    return math.sqrt(float(qb[p:m] @ qb[p:m])), iters


if __name__ == "__main__":
    fh = open(sys.argv[1], "rb")
    (count,) = struct.unpack("<i", fh.read(4))
    limit = int(sys.argv[2]) if len(sys.argv) > 2 else count
    bad = flips = 0
    worst = 0.0
    its = []
    for s in range(min(count, limit)):
        m, n = struct.unpack("<ii", fh.read(8))
        b = np.frombuffer(fh.read(8 * m))
        R = np.frombuffer(fh.read(8 * m * n)).reshape(n, m)
        (want,) = struct.unpack("<d", fh.read(8))
        if m > MR or n > MC:
            continue
        cols = [[(i, float(R[j, i])) for i in range(m) if R[j, i] != 0.0] for j in range(n)]
        r, it = emulate(m, n, b, cols)
        d = abs(r - want)
        worst = max(worst, d)
        if d > 1e-9 + 1e-8 * want:
            bad += 1
            print("system", s, "residual", r, "expected", want)
        flips += (r <= 1e-6) != (want <= 1e-6)
        its.append(it)
    print(f"{len(its)} systems: residual mismatches {bad} (worst |diff| {worst:.2e}), verdict flips {flips}, "
          f"passive-set solves mean {np.mean(its):.1f} max {max(its)}")
