"""Restatement of the compas 2.1.1 pieces the reference's `Shape` relies on.

TEST INFRASTRUCTURE (see oracle/__init__.py).  compas / compas_robots are not
vendored under /root/reference and are not installable here; this file restates
their published behaviour for exactly the calls made by

  * `Shape.from_urdf`   assembly_gym/assembly_gym/envs/assembly_env.py:54-68
  * `Shape.from_mesh`   assembly_gym/assembly_gym/envs/assembly_env.py:45-51
  * `merge_coplanar_faces` / `is_coplanar` / `contains_point`
                        assembly_gym/assembly_gym/utils/geometry.py:9-36

Pinned by: the merged hexagon face dict printed in
notebooks/CRA_Assembly.ipynb cells 24-25 (reproduced key-for-key and
cycle-for-cycle by `merge_coplanar_faces` below, see tests/test_oracle_shapes.py)
and the STL vertex order printed in cell 31.
"""
import math
import os
import struct
import xml.etree.ElementTree as ET


# ---------------------------------------------------------------- vectors
def subtract_vectors(a, b):
    return [a[0] - b[0], a[1] - b[1], a[2] - b[2]]


def cross_vectors(a, b):
    return [a[1] * b[2] - a[2] * b[1],
            a[2] * b[0] - a[0] * b[2],
            a[0] * b[1] - a[1] * b[0]]


def dot_vectors(a, b):
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]


def length_vector(a):
    return math.sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2])


def normalize_vector(a):
    ln = length_vector(a)
    if not ln:
        return list(a)
    return [a[0] / ln, a[1] / ln, a[2] / ln]


def centroid_points(points):
    """compas.geometry.centroid_points: component sums divided by the count."""
    p = len(points)
    sx = sy = sz = 0.0
    for q in points:
        sx += q[0]
        sy += q[1]
        sz += q[2]
    return [sx / p, sy / p, sz / p]


def normal_polygon(polygon, unitized=True):
    """compas.geometry.normal_polygon: sum of cross products about the centroid."""
    o = centroid_points(polygon)
    a = polygon[-1]
    oa = subtract_vectors(a, o)
    nx = ny = nz = 0.0
    for b in polygon:
        ob = subtract_vectors(b, o)
        n = cross_vectors(oa, ob)
        oa = ob
        nx += n[0]
        ny += n[1]
        nz += n[2]
    if not unitized:
        return [nx, ny, nz]
    return normalize_vector([nx, ny, nz])


# ---------------------------------------------------------------- mesh
class Mesh:
    """Half-edge mesh with compas' dict semantics (insertion-ordered faces,
    new faces keyed max+1)."""

    def __init__(self):
        self.vertex = {}
        self.face = {}
        self.halfedge = {}
        self._max_face = -1

    @classmethod
    def from_vertices_and_faces(cls, vertices, faces):
        mesh = cls()
        for key, xyz in enumerate(vertices):
            mesh.vertex[key] = [float(xyz[0]), float(xyz[1]), float(xyz[2])]
            mesh.halfedge[key] = {}
        for face in faces:
            mesh.add_face(list(face))
        return mesh

    def copy(self):
        other = Mesh()
        other.vertex = {k: list(v) for k, v in self.vertex.items()}
        other.face = {k: list(v) for k, v in self.face.items()}
        other.halfedge = {k: dict(v) for k, v in self.halfedge.items()}
        other._max_face = self._max_face
        return other

    # -- topology
    def add_face(self, vertices):
        if vertices[-1] == vertices[0]:
            vertices = vertices[:-1]
        self._max_face += 1
        fkey = self._max_face
        self.face[fkey] = vertices
        for u, v in zip(vertices, vertices[1:] + vertices[:1]):
            self.halfedge[u][v] = fkey
            if u not in self.halfedge[v]:
                self.halfedge[v][u] = None
        return fkey

    def delete_face(self, fkey):
        for u, v in self.face_halfedges(fkey):
            self.halfedge[u][v] = None
            if self.halfedge[v][u] is None:
                del self.halfedge[u][v]
                del self.halfedge[v][u]
        del self.face[fkey]

    def faces(self):
        return iter(list(self.face))

    def vertices(self):
        return iter(list(self.vertex))

    def face_halfedges(self, fkey):
        vs = self.face[fkey]
        return list(zip(vs, vs[1:] + vs[:1]))

    def face_neighbors(self, fkey):
        nbrs = []
        for u, v in self.face_halfedges(fkey):
            nbr = self.halfedge[v].get(u)
            if nbr is not None:
                nbrs.append(nbr)
        return nbrs

    def face_neighborhood(self, fkey):
        # ring=1; compas builds a python set and returns list(set) -- keep the
        # set so the iteration order is CPython's, as in the reference.
        nbrs = set(self.face_neighbors(fkey))
        nbrs.discard(fkey)
        return list(nbrs)

    def merge_faces(self, fkeys):
        """Merge two faces over their shared edge (compas Mesh.merge_faces).
        The new cycle starts at the head of the shared half-edge of the first
        face; this reproduces the cycles stored in CRA_Assembly.ipynb cell 24."""
        a, b = fkeys
        va, vb = self.face[a], self.face[b]
        for u, v in self.face_halfedges(a):
            if self.halfedge[v].get(u) == b:
                break
        else:
            raise ValueError("faces do not share an edge")
        ia = va.index(v)
        ib = vb.index(u)
        ra = va[ia:] + va[:ia]          # v ... u
        rb = vb[ib:] + vb[:ib]          # u ... v
        cycle = ra[:-1] + rb[:-1]
        self.delete_face(a)
        self.delete_face(b)
        return self.add_face(cycle)

    # -- geometry
    def vertex_coordinates(self, key):
        return list(self.vertex[key])

    def face_coordinates(self, fkey):
        return [self.vertex_coordinates(k) for k in self.face[fkey]]

    def face_normal(self, fkey, unitized=True):
        return normal_polygon(self.face_coordinates(fkey), unitized=unitized)

    def face_center(self, fkey):
        return centroid_points(self.face_coordinates(fkey))

    def aabb(self):
        xs, ys, zs = zip(*self.vertex.values())
        return (min(xs), min(ys), min(zs)), (max(xs), max(ys), max(zs))


# ---------------------------------------------------------------- geometry.py
def contains_point(plane_point, plane_normal, point, tol=1e-6):
    """geometry.py:24-26."""
    return abs(dot_vectors(subtract_vectors(point, plane_point), plane_normal)) <= tol


def is_coplanar(points):
    """geometry.py:29-36 (Plane.from_three_points = unit normal of the first three)."""
    if len(points) < 4:
        return True
    a, b, c = points[:3]
    normal = normalize_vector(cross_vectors(subtract_vectors(b, a), subtract_vectors(c, a)))
    for p in points[3:]:
        if not contains_point(a, normal, p):
            return False
    return True


def merge_coplanar_faces(mesh):
    """geometry.py:9-21."""
    faces = [*mesh.faces()]
    while len(faces) > 0:
        face = faces.pop()
        for face2 in mesh.face_neighborhood(face):
            points = mesh.face_coordinates(face) + mesh.face_coordinates(face2)
            if is_coplanar(points):
                new_face = mesh.merge_faces([face, face2])
                faces.remove(face2)
                faces.append(new_face)
                break


# ---------------------------------------------------------------- file formats
def _geometric_key(xyz, precision=3):
    # compas TOL.geometric_key with the default precision (3 decimals), -0.0 -> 0.0
    fmt = "{{0:.{0}f}}".format(precision)
    out = []
    for c in xyz:
        s = fmt.format(c)
        if float(s) == 0.0:
            s = fmt.format(0.0)
        out.append(s)
    return ",".join(out)


def read_stl_facets(path):
    """Binary or ASCII STL -> list of facets, each three (x, y, z) tuples.
    Binary coordinates are IEEE float32 widened to float64 (struct 'f')."""
    with open(path, "rb") as fh:
        data = fh.read()
    is_ascii = data[:5] == b"solid" and b"facet" in data[:512]
    facets = []
    if is_ascii:
        pts = []
        for line in data.decode("ascii", "replace").splitlines():
            parts = line.split()
            if parts and parts[0] == "vertex":
                pts.append(tuple(float(x) for x in parts[1:4]))
                if len(pts) == 3:
                    facets.append(pts)
                    pts = []
        return facets
    (count,) = struct.unpack_from("<I", data, 80)
    for i in range(count):
        rec = struct.unpack_from("<12fH", data, 84 + 50 * i)
        facets.append([tuple(rec[3:6]), tuple(rec[6:9]), tuple(rec[9:12])])
    return facets


def mesh_from_stl(path):
    """compas Mesh.from_stl: vertices welded by geometric key in first-appearance
    order (CRA_Assembly.ipynb cell 31 shows that order for the trapezoid)."""
    index = {}
    vertices = []
    faces = []
    for facet in read_stl_facets(path):
        face = []
        for xyz in facet:
            key = _geometric_key(xyz)
            if key not in index:
                index[key] = len(vertices)
                vertices.append(xyz)
            face.append(index[key])
        faces.append(face)
    return Mesh.from_vertices_and_faces(vertices, faces)


def mesh_from_box(xsize, ysize, zsize):
    """compas 2.1.1 Box.to_vertices_and_faces for a box centred on the world frame:
    vertices a..h and faces bottom, front(-y), right(+x), back(+y), left(-x), top."""
    hx, hy, hz = 0.5 * xsize, 0.5 * ysize, 0.5 * zsize
    a = [-hx, -hy, -hz]
    b = [-hx, +hy, -hz]
    c = [+hx, +hy, -hz]
    d = [+hx, -hy, -hz]
    e = [a[0], a[1], a[2] + zsize]
    f = [d[0], d[1], d[2] + zsize]
    g = [c[0], c[1], c[2] + zsize]
    h = [b[0], b[1], b[2] + zsize]
    faces = [[0, 1, 2, 3], [0, 3, 5, 4], [3, 2, 6, 5], [2, 1, 7, 6], [1, 0, 4, 7], [4, 5, 6, 7]]
    return Mesh.from_vertices_and_faces([a, b, c, d, e, f, g, h], faces)


def mesh_from_urdf(urdf_file, package="blocks"):
    """`robot.links[0].collision[0].geometry.shape.meshes[0]`
    (assembly_env.py:65-67) for the two geometry kinds the block library uses."""
    root = ET.parse(urdf_file).getroot()
    link = root.findall("link")[0]
    geometry = link.findall("collision")[0].find("geometry")
    box = geometry.find("box")
    if box is not None:
        sx, sy, sz = (float(x) for x in box.get("size").split())
        return mesh_from_box(sx, sy, sz)
    mesh = geometry.find("mesh")
    if mesh is None:
        raise ValueError("unsupported collision geometry in %s" % urdf_file)
    filename = mesh.get("filename")
    prefix = "package://%s/" % package
    if not filename.startswith(prefix):
        raise ValueError("unsupported mesh url %s" % filename)
    base_path = os.path.split(urdf_file)[0]
    return mesh_from_stl(os.path.join(base_path, package, filename[len(prefix):]))
