"""Oracle restatement of the stability path.

TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows
  * `AssemblyEnv._reset_cra_assembly`  assembly_env.py:281-304  (floor support
    Box(width=10, depth=10, thickness=0.5) centred at (0, 0, -0.25), node -1 fixed;
    blocks 0..n-1; `assembly_interfaces_numpy(assembly, amin=0.001)`),
  * `is_stable_rbe`                    utils/stability.py:49-71 (edge-less rule,
    "infeasible" -> False, other solver errors -> None),
  * compas_cra `assembly_interfaces_numpy` and `rbe_solve(penalty=False)`, which are
    NOT under /root/reference (requirements: git+https://github.com/kirschnj/compas_cra,
    unpinned; docker/cscs/requirements.txt:5-6).  Their published algorithm is
    restated in its exact 2-D reduction (all bodies are prisms symmetric about
    y = 0): face-face interfaces between coplanar, opposed
    faces with overlap area >= amin; per contact point a normal force fn >= 0
    and a tangential force |ft| <= mu*fn; three equilibrium rows per free block.

PARITY UNPINNED for: tmax, contact-force values, objective of the force QP.
Pinned for verdicts by utils/structures.py:22-108 labels and the notebook runs
(tests/test_oracle_fixtures.py).
"""
import math

import numpy as np

FLOOR = -1


class Interface:
    """One face-face contact: bodies a < b (node order), contact normal = outward
    normal of a's face, two contact points (ends of the overlap segment)."""
    __slots__ = ("a", "b", "face_a", "face_b", "normal", "tangent", "points", "size")

    def __init__(self, a, b, face_a, face_b, normal, tangent, points, size):
        self.a, self.b, self.face_a, self.face_b = a, b, face_a, face_b
        self.normal, self.tangent, self.points, self.size = normal, tangent, points, size


class _Body:
    __slots__ = ("node", "normals", "centers", "ends", "depth", "com", "area", "is_support", "radius")


def _floor_body(bounds):
    """assembly_env.py:290-296: Box(width, depth, 0.05*width) centred at x = y = 0
    with its top face on z = 0.  Only the top face can touch a block."""
    width = bounds[1][0] - bounds[0][0]
    depth = bounds[1][1] - bounds[0][1]
    body = _Body()
    body.node = FLOOR
    body.normals = [(0.0, 1.0)]
    body.centers = [(0.0, 0.0)]
    body.ends = [((-0.5 * width, 0.0), (0.5 * width, 0.0))]
    body.depth = depth
    body.com = (0.0, -0.025 * width)
    body.area = width * 0.05 * width
    body.is_support = True
    body.radius = 0.0
    return body


def _block_body(node, block):
    body = _Body()
    body.node = node
    body.normals = block.face_normals_2d
    body.centers = block.face_centers_2d
    body.ends = block.face_ends_2d
    body.depth = block.depth
    body.com = block.centroid_2d
    body.area = block.area
    body.is_support = bool(block.is_static)
    body.radius = block.radius
    return body


def find_interfaces(bodies, tmax=1e-6, amin=1e-3):
    """2-D restatement of compas_cra `assembly_interfaces_numpy` (all pairs)."""
    out = []
    for ia in range(len(bodies)):
        A = bodies[ia]
        for ib in range(ia + 1, len(bodies)):
            B = bodies[ib]
            dmin = min(A.depth, B.depth)
            for fa in range(len(A.normals)):
                nx, nz = A.normals[fa]
                cx, cz = A.centers[fa]
                tx, tz = nz, -nx                       # face frame x axis (n_z, -n_x)
                (a0x, a0z), (a1x, a1z) = A.ends[fa]
                sa0 = (a0x - cx) * tx + (a0z - cz) * tz
                sa1 = (a1x - cx) * tx + (a1z - cz) * tz
                alo, ahi = min(sa0, sa1), max(sa0, sa1)
                for fb in range(len(B.normals)):
                    mx, mz = B.normals[fb]
                    if nx * mx + nz * mz >= 0.0:       # faces must oppose each other
                        continue
                    (b0x, b0z), (b1x, b1z) = B.ends[fb]
                    d0 = (b0x - cx) * nx + (b0z - cz) * nz
                    d1 = (b1x - cx) * nx + (b1z - cz) * nz
                    if abs(d0) > tmax or abs(d1) > tmax:   # coplanarity in a's face frame
                        continue
                    sb0 = (b0x - cx) * tx + (b0z - cz) * tz
                    sb1 = (b1x - cx) * tx + (b1z - cz) * tz
                    lo = max(alo, min(sb0, sb1))
                    hi = min(ahi, max(sb0, sb1))
                    size = (hi - lo) * dmin
                    if not size >= amin:
                        continue
                    p0 = (cx + lo * tx, cz + lo * tz)
                    p1 = (cx + hi * tx, cz + hi * tz)
                    out.append(Interface(A.node, B.node, fa, fb, (nx, nz), (tx, tz), (p0, p1), size))
    return out


class CRAAssembly:
    """Stand-in for compas_cra's CRA_Assembly as the reference uses it."""

    def __init__(self, bounds, blocks, tmax=1e-6, amin=1e-3):
        self.bodies = [_floor_body(bounds)] + [_block_body(i, b) for i, b in enumerate(blocks)]
        self.interfaces = find_interfaces(self.bodies, tmax, amin) if len(blocks) > 0 else []
        self.forces = None

    def set_boundary_condition(self, node):
        self.bodies[node + 1].is_support = True

    def number_of_edges(self):
        return len({(i.a, i.b) for i in self.interfaces})

    def free_nodes(self):
        return [b.node for b in self.bodies if not b.is_support]


def equilibrium_system(assembly, mu, density):
    """A f = b with f = (fn_0, ft_0, fn_1, ft_1, ...) over the contact points of all
    interfaces.  Force on body b of an interface: fn*n + ft*t, on body a the opposite.
    Rows per free block j: sum Fx = 0, sum Fz = W_j, sum (p - com_j) x F / L0 = 0,
    W_j = density * area_j * depth_j  (weight acts along -z).  L0, the largest
    centre-of-mass -> vertex distance among the blocks, makes the torque rows
    commensurable with the force rows (it only matters for the residual r*)."""
    free = assembly.free_nodes()
    row_of = {node: 3 * k for k, node in enumerate(free)}
    ncp = 2 * len(assembly.interfaces)
    L0 = max([body.radius for body in assembly.bodies] + [1e-300])
    A = np.zeros((3 * len(free), 2 * ncp))
    b = np.zeros(3 * len(free))
    for node in free:
        body = assembly.bodies[node + 1]
        b[row_of[node] + 1] = density * body.area * body.depth
    col = 0
    for itf in assembly.interfaces:
        nx, nz = itf.normal
        tx, tz = itf.tangent
        for (px, pz) in itf.points:
            for node, sign in ((itf.a, -1.0), (itf.b, 1.0)):
                if node in row_of:
                    r = row_of[node]
                    gx, gz = assembly.bodies[node + 1].com
                    rx, rz = px - gx, pz - gz
                    A[r + 0, col] = sign * nx
                    A[r + 1, col] = sign * nz
                    A[r + 2, col] = sign * (rx * nz - rz * nx) / L0
                    A[r + 0, col + 1] = sign * tx
                    A[r + 1, col + 1] = sign * tz
                    A[r + 2, col + 1] = sign * (rx * tz - rz * tx) / L0
            col += 2
    return A, b


def rbe_feasible(A, b, mu):
    """LP feasibility of  A f = b, fn >= 0, |ft| <= mu*fn  with HiGHS.
    Returns True / False / None (solver error), like rbe_solve's outcome mapping
    in utils/stability.py:59-68.  HiGHS occasionally answers "unknown" on
    infeasible instances; the simplex and interior-point variants are tried
    in turn before giving up."""
    from scipy.optimize import linprog
    m, n = A.shape
    ncp = n // 2
    if m == 0:
        return True
    if n == 0:
        return bool(np.all(b == 0))
    # friction rows: +-ft - mu*fn <= 0
    Aub = np.zeros((2 * ncp, n))
    for k in range(ncp):
        Aub[2 * k, 2 * k] = -mu
        Aub[2 * k, 2 * k + 1] = 1.0
        Aub[2 * k + 1, 2 * k] = -mu
        Aub[2 * k + 1, 2 * k + 1] = -1.0
    bounds = [(0, None), (None, None)] * ncp
    for method in ("highs", "highs-ds", "highs-ipm"):
        res = linprog(np.zeros(n), A_ub=Aub, b_ub=np.zeros(2 * ncp), A_eq=A, b_eq=b, bounds=bounds, method=method)
        if res.status == 0:
            return True
        if res.status == 2:
            return False
    return None


def is_stable_rbe(assembly_env):
    """utils/stability.py:49-71."""
    asm = assembly_env.cra_assembly
    if asm.number_of_edges() == 0:
        return len(asm.free_nodes()) == 0, None
    A, b = equilibrium_system(asm, assembly_env.mu, assembly_env.density)
    res = rbe_feasible(A, b, assembly_env.mu)
    if res is None:
        return None, dict(error="solver")
    return res, None


# ---------------------------------------------------------------- margin / forces
def ray_matrix(A, mu):
    """Columns A(n + mu t), A(n - mu t): the friction cone's two edge rays."""
    ncp = A.shape[1] // 2
    R = np.zeros((A.shape[0], 2 * ncp))
    R[:, 0::2] = A[:, 0::2] + mu * A[:, 1::2]
    R[:, 1::2] = A[:, 0::2] - mu * A[:, 1::2]
    return R


def equilibrium_residual(A, b, mu):
    """r* = min over the friction cones of ||A f - b|| / ||b||: the verdict margin
    of SURVEY.md section 8(d).  Bounded-variable least squares on the ray form
    (scipy lsq_linear/BVLS, an active-set method with exact KKT termination);
    the KKT conditions are re-checked here because scipy.optimize.nnls was seen
    to return wrong minimisers on these systems."""
    from scipy.optimize import lsq_linear
    if A.shape[0] == 0:
        return 0.0
    nb = float(np.linalg.norm(b))
    if nb == 0.0:
        return 0.0
    if A.shape[1] == 0:
        return 1.0
    R = ray_matrix(A, mu)
    res = lsq_linear(R, b, bounds=(0, np.inf), method="bvls", tol=1e-14, max_iter=20 * R.shape[1] + 100)
    x = res.x
    resid = b - R @ x
    w = R.T @ resid                          # must be <= 0 on x = 0 and == 0 on x > 0
    scale = max(1.0, float(np.abs(R).max())) * nb
    free = x > 1e-12 * max(1.0, float(np.abs(x).max()))
    if np.any(w[~free] > 1e-6 * scale) or np.any(np.abs(w[free]) > 1e-6 * scale):
        raise RuntimeError("BVLS did not reach a KKT point")
    return float(np.linalg.norm(resid)) / nb


RHO_SCHEDULE = (1e4, 1e8, 1e8, 1e8, 1e8, 1e8)


def min_norm_forces(A, b, mu, schedule=RHO_SCHEDULE, max_newton=60):
    """argmin ||f||^2  s.t.  A f = b, f in the friction cones  -- the canonical
    force definition of this build (the reference's IPOPT objective is unpinned,
    SURVEY.md section 8c).  Proximal-point iteration on the concave dual

        d(y) = b.y - 1/2 ||P_K(A^T y)||^2,      f = P_K(A^T y),

    each proximal step  max_y d(y) - ||y - y_k||^2 / (2 rho)  solved by a
    semismooth Newton method with a derivative-based line search.  For an
    infeasible system the residual b - A f converges to the minimum-norm
    residual r*.  Returns (f, y, relative residual, status) with status in
    {"feasible", "stagnated", "maxouter"}."""
    m, n = A.shape
    if m == 0 or n == 0:
        return np.zeros(n), np.zeros(m), (0.0 if m == 0 or not np.any(b) else 1.0), "feasible"
    nb = float(np.linalg.norm(b))
    if nb == 0.0:
        return np.zeros(n), np.zeros(m), 0.0, "feasible"
    bs = b / nb
    y = np.zeros(m)
    first = True
    rprev = None
    status = "maxouter"
    r = 1.0
    for rho in schedule:
        yk = y.copy()
        for _ in range(max_newton):
            g = A.T @ y
            f, J = _project_cones(g, mu)
            if first:
                J[:] = np.eye(2)
                first = False
            grad = bs - A @ f - (y - yk) / rho
            if np.linalg.norm(grad) <= 1e-10:
                break
            H = _AJAt(A, J)
            H[np.diag_indices(m)] += 1.0 / rho
            d = np.linalg.solve(H, grad)
            h = A.T @ d
            dd, bd, yd = float(d @ d), float(bs @ d), float((y - yk) @ d)
            phi0 = float(grad @ d)
            if phi0 <= 1e-30:
                break

            def dphi(t):
                ft, _ = _project_cones(g + t * h, mu)
                return bd - float(ft @ h) - (yd + t * dd) / rho

            t = 1.0
            p = dphi(t)
            if p < -0.1 * phi0:          # full step kept when it passes the search's own acceptance test
                # bracket the root of the piecewise-linear derivative (safeguarded regula falsi)
                lo, plo, hi, phi = 0.0, phi0, 1.0, p
                for _ls in range(20):
                    w = hi - lo
                    t = lo + w * plo / (plo - phi)
                    t = min(max(t, lo + 0.1 * w), hi - 0.1 * w)
                    p = dphi(t)
                    if abs(p) <= 0.1 * phi0:
                        break
                    if p > 0.0:
                        lo, plo = t, p
                    else:
                        hi, phi = t, p
                if p < 0.0 and abs(p) > 0.1 * phi0 and lo > 0.0:
                    t = lo
            y = y + t * d
            if t * math.sqrt(dd) <= 1e-15 * max(1.0, float(np.linalg.norm(y))):
                break
        f, _ = _project_cones(A.T @ y, mu)
        r = float(np.linalg.norm(bs - A @ f))
        if r <= 1e-9:
            status = "feasible"
            break
        if rprev is not None and abs(r - rprev) <= 1e-3 * r:
            status = "stagnated"
            break
        if rprev is not None and r >= 0.9 * rprev and r > 1e-3:
            status = "stagnated"            # stalled far above the verdict threshold
            break
        rprev = r
    return f * nb, y * nb, r, status


def _project_cones(g, mu):
    """Euclidean projection of (gn, gt) pairs onto {|ft| <= mu fn}; returns the
    projection and, per contact point, the 2x2 generalised Jacobian."""
    ncp = g.size // 2
    f = np.zeros_like(g)
    J = np.zeros((ncp, 2, 2))
    den = 1.0 + mu * mu
    for k in range(ncp):
        gn, gt = g[2 * k], g[2 * k + 1]
        if abs(gt) <= mu * gn:
            f[2 * k], f[2 * k + 1] = gn, gt
            J[k] = np.eye(2)
        elif mu * abs(gt) <= -gn:
            pass
        else:
            sg = 1.0 if gt > 0 else -1.0
            kk = (gn + mu * abs(gt)) / den
            f[2 * k], f[2 * k + 1] = kk, sg * mu * kk
            u = np.array([1.0, sg * mu])
            J[k] = np.outer(u, u) / den
    return f, J


def _AJAt(A, J):
    m = A.shape[0]
    H = np.zeros((m, m))
    for k in range(J.shape[0]):
        G = A[:, 2 * k:2 * k + 2]
        H += G @ J[k] @ G.T
    return H
