"""Block library loader: URDF + STL -> the 2-D face tables the kernels consume.

Host-side mirror of `Shape.from_urdf` / `Shape.from_mesh`
(assembly_gym/assembly_gym/envs/assembly_env.py:45-68) including the face order produced
by `merge_coplanar_faces` (assembly_gym/assembly_gym/utils/geometry.py:9-21) on a compas
mesh: that order IS the `face` index of the Action API, so it is reproduced exactly
(same vertex welding, same neighbour iteration, same key numbering).
"""
import math
import os
import struct
import xml.etree.ElementTree as ET

PACKAGE_ROOT = os.path.dirname(os.path.abspath(__file__))


def resolve_urdf(urdf_file):
    """assembly_env.py:55-60: as given, else relative to the package that ships `shapes/`."""
    if os.path.exists(urdf_file):
        return urdf_file
    for root in (os.environ.get("BRIDGES_B200_SHAPES_ROOT"), PACKAGE_ROOT):
        if root:
            cand = os.path.join(root, urdf_file)
            if os.path.exists(cand):
                return cand
    raise FileNotFoundError(f"URDF file not found: {urdf_file}")


# ---------------------------------------------------------------- triangle soup
def _stl_triangles(path):
    with open(path, "rb") as fh:
        raw = fh.read()
    if raw[:5] == b"solid" and b"facet" in raw[:512]:
        nums = [tuple(float(t) for t in ln.split()[1:4])
                for ln in raw.decode("ascii", "replace").splitlines() if ln.strip().startswith("vertex")]
        return [nums[i:i + 3] for i in range(0, len(nums), 3)]
    n = struct.unpack_from("<I", raw, 80)[0]
    tris = []
    for k in range(n):
        v = struct.unpack_from("<9f", raw, 84 + 50 * k + 12)
        tris.append([v[0:3], v[3:6], v[6:9]])
    return tris


def _weld(tris):
    """Vertices in first-appearance order, welded at 3 decimals (compas geometric key)."""
    seen, verts, faces = {}, [], []
    for tri in tris:
        ids = []
        for p in tri:
            key = tuple("%.3f" % (c + 0.0 if round(c, 3) != 0 else 0.0) for c in p)
            if key not in seen:
                seen[key] = len(verts)
                verts.append([float(p[0]), float(p[1]), float(p[2])])
            ids.append(seen[key])
        faces.append(ids)
    return verts, faces


def _box(sx, sy, sz):
    hx, hy, hz = 0.5 * sx, 0.5 * sy, 0.5 * sz
    lo = [[-hx, -hy, -hz], [-hx, hy, -hz], [hx, hy, -hz], [hx, -hy, -hz]]
    a, b, c, d = lo
    verts = lo + [[a[0], a[1], a[2] + sz], [d[0], d[1], d[2] + sz], [c[0], c[1], c[2] + sz], [b[0], b[1], b[2] + sz]]
    faces = [[0, 1, 2, 3], [0, 3, 5, 4], [3, 2, 6, 5], [2, 1, 7, 6], [1, 0, 4, 7], [4, 5, 6, 7]]
    return verts, faces


class _PolyMesh:
    """Faces keyed by increasing integers (dict order = creation order) plus a directed-edge
    -> face map; just enough to replay the reference's coplanar merge."""

    def __init__(self, verts, faces):
        self.v = verts
        self.f = {}
        self.edge = {}
        self.next_key = 0
        for cyc in faces:
            self.add(list(cyc))

    def add(self, cyc):
        key = self.next_key
        self.next_key += 1
        self.f[key] = cyc
        for u, w in zip(cyc, cyc[1:] + cyc[:1]):
            self.edge[(u, w)] = key
        return key

    def drop(self, key):
        cyc = self.f.pop(key)
        for u, w in zip(cyc, cyc[1:] + cyc[:1]):
            del self.edge[(u, w)]

    def neighbours(self, key):
        cyc = self.f[key]
        out = set()
        for u, w in zip(cyc, cyc[1:] + cyc[:1]):
            other = self.edge.get((w, u))
            if other is not None and other != key:
                out.add(other)
        return list(out)            # CPython set order, as compas' face_neighborhood

    def coords(self, key):
        return [self.v[i] for i in self.f[key]]

    def join(self, a, b):
        ca, cb = self.f[a], self.f[b]
        for u, w in zip(ca, ca[1:] + ca[:1]):
            if self.edge.get((w, u)) == b:
                break
        ia, ib = ca.index(w), cb.index(u)
        cyc = (ca[ia:] + ca[:ia])[:-1] + (cb[ib:] + cb[:ib])[:-1]
        self.drop(a)
        self.drop(b)
        return self.add(cyc)


def _coplanar(pts, tol=1e-6):
    if len(pts) < 4:
        return True
    (ax, ay, az), (bx, by, bz), (cx, cy, cz) = pts[:3]
    ux, uy, uz, vx, vy, vz = bx - ax, by - ay, bz - az, cx - ax, cy - ay, cz - az
    nx, ny, nz = uy * vz - uz * vy, uz * vx - ux * vz, ux * vy - uy * vx
    ln = math.sqrt(nx * nx + ny * ny + nz * nz)
    nx, ny, nz = nx / ln, ny / ln, nz / ln
    return all(abs((p[0] - ax) * nx + (p[1] - ay) * ny + (p[2] - az) * nz) <= tol for p in pts[3:])


def _merge_coplanar(mesh):
    work = list(mesh.f)
    while work:
        key = work.pop()
        for other in mesh.neighbours(key):
            if _coplanar(mesh.coords(key) + mesh.coords(other)):
                merged = mesh.join(key, other)
                work.remove(other)
                work.append(merged)
                break


def _centroid(pts):
    n = len(pts)
    sx = sy = sz = 0.0
    for p in pts:
        sx += p[0]
        sy += p[1]
        sz += p[2]
    return [sx / n, sy / n, sz / n]


def _unit_normal(pts):
    """compas normal_polygon: summed cross products about the vertex centroid, unitised."""
    o = _centroid(pts)
    px, py, pz = pts[-1][0] - o[0], pts[-1][1] - o[1], pts[-1][2] - o[2]
    nx = ny = nz = 0.0
    for q in pts:
        qx, qy, qz = q[0] - o[0], q[1] - o[1], q[2] - o[2]
        nx += py * qz - pz * qy
        ny += pz * qx - px * qz
        nz += px * qy - py * qx
        px, py, pz = qx, qy, qz
    ln = math.sqrt(nx * nx + ny * ny + nz * nz)
    return [nx / ln, ny / ln, nz / ln]


class ShapeTables:
    """Everything `bw_shape_desc` needs, plus the 3-D vertices for `Shape.vertices`."""
    __slots__ = ("urdf_file", "vertices3d", "face_keys", "face_cycles", "normals", "centers", "ends",
                 "polygon", "com", "area", "depth", "ymin", "ymax", "aabb")


def load_shape_tables(urdf_file, package="blocks"):
    path = resolve_urdf(urdf_file)
    link = ET.parse(path).getroot().findall("link")[0]
    geom = link.findall("collision")[0].find("geometry")
    if geom.find("box") is not None:
        verts, faces = _box(*(float(t) for t in geom.find("box").get("size").split()))
    else:
        name = geom.find("mesh").get("filename")
        prefix = "package://%s/" % package
        if not name.startswith(prefix):
            raise ValueError(f"unsupported mesh url {name}")
        verts, faces = _weld(_stl_triangles(os.path.join(os.path.dirname(path), package, name[len(prefix):])))
    mesh = _PolyMesh(verts, faces)
    _merge_coplanar(mesh)

    t = ShapeTables()
    t.urdf_file = path
    t.vertices3d = [list(v) for v in verts]
    normals3 = {k: _unit_normal(mesh.coords(k)) for k in mesh.f}
    keys2d = [k for k in mesh.f if abs(normals3[k][1]) < 1e-6]
    t.face_keys = keys2d + [k for k in mesh.f if k not in keys2d]
    t.face_cycles = {k: list(mesh.f[k]) for k in mesh.f}
    t.normals, t.centers, t.ends = [], [], []
    for k in keys2d:
        c = _centroid(mesh.coords(k))
        t.normals.append((normals3[k][0], normals3[k][2]))
        t.centers.append((c[0], c[2]))
        uniq = []
        for p in mesh.coords(k):
            if (p[0], p[2]) not in uniq:
                uniq.append((p[0], p[2]))
        if len(uniq) != 2:
            raise ValueError("block is not a prism along y")
        t.ends.append((uniq[0], uniq[1]))
    front = None
    for k in mesh.f:                       # Shape.vertices_2d: first face looking along +y
        front = k
        if abs(normals3[k][1] - 1) < 1e-3:
            break
    t.polygon = [(mesh.v[i][0], mesh.v[i][2]) for i in mesh.f[front]]
    ys = [v[1] for v in verts]
    t.ymin, t.ymax = min(ys), max(ys)
    t.depth = t.ymax - t.ymin
    xs = [v[0] for v in verts]
    zs = [v[2] for v in verts]
    t.aabb = ((min(xs), t.ymin, min(zs)), (max(xs), t.ymax, max(zs)))
    a2 = cx = cz = 0.0
    for i, (xa, za) in enumerate(t.polygon):
        xb, zb = t.polygon[(i + 1) % len(t.polygon)]
        w = xa * zb - xb * za
        a2 += w
        cx += (xa + xb) * w
        cz += (za + zb) * w
    t.area = abs(a2) / 2.0
    t.com = (cx / (3.0 * a2), cz / (3.0 * a2))
    return t
