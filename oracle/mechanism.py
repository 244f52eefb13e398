"""Rigid-mechanism certificates of "no equilibrium" -- CPU restatement of `Solver::screen`
(bridges_b200 csrc/bw_solver.cuh), the shortcut the CUDA step takes before its equilibrium solve.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference has no counterpart: `is_stable_rbe`
(utils/stability.py:49-71) always builds and solves the full problem.  The certificate is a Farkas
vector of that problem, so it can only confirm the verdict "infeasible" the reference's solver would
reach (tests/test_oracle_mechanism.py checks it against HiGHS).

Mathematics.  Let S be a set of free blocks moved by one rigid virtual motion n = (ux, uz, w L0)
(translation + rotation about the origin).  Contacts inside S see no relative motion; a contact point
p between S and the rest contributes its two friction-cone edge rays r+- = (F, (p x F) / L0),
F = n_c +- mu t_c, signed by sigma = +1 when S holds body b of the interface (the body the contact
normal points into) and -1 when it holds body a.  If

    n . (sigma r) >= 0  for every boundary ray     and     n . b_S < 0,

with b_S = (0, W_S, sum_j x_j W_j / L0) the wrench the contact forces on S have to deliver, then no
non-negative combination of the rays equals b_S: A f = b has no solution in the friction cones.

Candidates (same as the kernel): S_i = block i plus every later free block touching a member, for
i from the last free block down; n = rotation about a boundary contact point (r+ x r-) or the
translations perpendicular to its two rays, both signs.  Margins EPS on the rays, DELTA on the weight.
"""
import numpy as np

EPS, DELTA = 1e-10, 1e-5


def _rays(assembly, mu):
    L0 = max([body.radius for body in assembly.bodies] + [1e-300])
    rays = []                                  # per contact point: (a, b, r_plus, r_minus), unit 3-vectors
    for itf in assembly.interfaces:
        nx, nz = itf.normal
        tx, tz = itf.tangent
        for (px, pz) in itf.points:
            wn = np.array([nx, nz, (px * nz - pz * nx) / L0])
            wt = np.array([tx, tz, (px * tz - pz * tx) / L0])
            rp, rm = wn + mu * wt, wn - mu * wt
            rays.append((itf.a, itf.b, rp / np.linalg.norm(rp), rm / np.linalg.norm(rm)))
    return rays, L0


def support_closures(assembly):
    """[(i, S_i)] for the free nodes i in descending order; S_i as a sorted list of nodes."""
    free = assembly.free_nodes()
    adj = {n: set() for n in free}
    for itf in assembly.interfaces:
        if itf.a in adj and itf.b in adj:
            adj[itf.a].add(itf.b)
            adj[itf.b].add(itf.a)
    out = []
    for k in range(len(free) - 1, -1, -1):
        S = {free[k]}
        for j in free[k + 1:]:
            if adj[j] & S:
                S.add(j)
        out.append((free[k], sorted(S)))
    return out


def mechanism_certificate(assembly, mu, density=1.0):
    """None, or dict(nodes=S, motion=n) proving that the assembly has no equilibrium."""
    free = assembly.free_nodes()
    if not free:
        return None
    rays, L0 = _rays(assembly, mu)
    weight = {n: density * assembly.bodies[n + 1].area * assembly.bodies[n + 1].depth for n in free}
    wsum = np.sqrt(sum(w * w for w in weight.values()))
    for _, S in support_closures(assembly):
        inS = set(S)
        bS = np.zeros(3)
        for n in S:
            w = weight[n] / wsum
            bS += (0.0, w, assembly.bodies[n + 1].com[0] * w / L0)
        boundary = []                          # signed rays
        for a, b, rp, rm in rays:
            if (b in inS) != (a in inS):
                sg = 1.0 if b in inS else -1.0
                boundary.append((sg * rp, sg * rm))
        if not boundary:
            if bS[1] > 0.0:
                return dict(nodes=S, motion=np.array([0.0, -1.0, 0.0]))
            continue
        bn = np.linalg.norm(bS)
        R = np.array([r for pair in boundary for r in pair])
        for rp, rm in boundary:
            for n in (np.cross(rp, rm), np.array([rp[1], -rp[0], 0.0]), np.array([rm[1], -rm[0], 0.0])):
                ln = np.linalg.norm(n)
                if ln <= 1e-9:
                    continue
                d = R @ n
                w = float(n @ bS)
                if d.min() >= -EPS * ln and w <= -DELTA * ln * bn:
                    return dict(nodes=S, motion=n / ln)
                if d.max() <= EPS * ln and w >= DELTA * ln * bn:
                    return dict(nodes=S, motion=-n / ln)
    return None
