"""Oracle restatement of assembly_gym/assembly_gym/envs/assembly_env.py.

TEST INFRASTRUCTURE (see oracle/__init__.py).

`Shape`  follows assembly_env.py:21-137, `Block` :140-156, `AssemblyEnv` :159-438
(default configuration: cra_env=True, pybullet_env=False).  All blocks of the
library are prisms extruded along y and every placement keeps y = 0
(gym_env.py:213), so a posed block is stored in the CANONICAL 2-D form

    pose = (tx, tz, c, s),   R(x, z) = (c*x + s*z,  c*z - s*x),   v' = R v + t

(rotation about +y by the angle whose cosine/sine are c/s), every product and
sum individually rounded in float64, no fused multiply-add.  The posed face
table is  n' = R n,  centre' = R centre + t  with n/centre the shape's compas
face normal / face centre.  The CUDA path implements the same sequence of
operations, which is what makes rasters and flags comparable bit for bit.
"""
import math
import os

from . import compas_lite as cl

_HERE = os.path.dirname(os.path.abspath(__file__))
# the block library that ships with the product package (same files as the
# reference's assembly_gym/shapes, see tests/test_oracle_shapes.py)
DEFAULT_SHAPES_ROOT = os.path.join(os.path.dirname(_HERE), "bridges-with-reinforcement-learning_b200")


def rot(c, s, x, z):
    """Canonical 2-D rotation (about +y): separate multiplies, then one add/sub."""
    return c * x + s * z, c * z - s * x


class Shape:
    """assembly_env.py:21-137."""

    def __init__(self, mesh=None, urdf_file=None, name="", receiving_faces_2d=None, target_faces_2d=None):
        self.urdf_file = None
        self.mesh = None
        self.name = name
        if mesh is not None:
            self.from_mesh(mesh)
        elif urdf_file is not None:
            self.from_urdf(urdf_file)
        self._target_faces_2d = target_faces_2d
        self._receiving_faces_2d = receiving_faces_2d

    # -- loading
    def from_mesh(self, mesh, merge_faces=True):
        if merge_faces:
            cl.merge_coplanar_faces(mesh)
        self.mesh = mesh.copy()
        self.bounding_box = mesh.aabb()
        self._2d_faces = [f for f in mesh.faces() if abs(mesh.face_normal(f)[1]) < 1e-6]
        self._all_faces = self._2d_faces + [f for f in mesh.faces() if f not in self._2d_faces]
        self._build_tables()

    def from_urdf(self, urdf_file, package="blocks", merge_faces=True):
        if not os.path.exists(urdf_file):
            urdf_file = os.path.join(os.environ.get("ORACLE_SHAPES_ROOT", DEFAULT_SHAPES_ROOT), urdf_file)
            if not os.path.exists(urdf_file):
                raise FileNotFoundError(f"URDF file not found: {urdf_file}")
        self.urdf_file = urdf_file
        self.from_mesh(cl.mesh_from_urdf(urdf_file, package), merge_faces=merge_faces)

    def _build_tables(self):
        mesh = self.mesh
        # 2-D face table in Action face-index order
        self.face_normals_2d = []
        self.face_centers_2d = []
        self.face_ends_2d = []
        for f in self._2d_faces:
            n = mesh.face_normal(f)
            c = mesh.face_center(f)
            self.face_normals_2d.append((n[0], n[2]))
            self.face_centers_2d.append((c[0], c[2]))
            ends = []
            for p in mesh.face_coordinates(f):
                q = (p[0], p[2])
                if q not in ends:
                    ends.append(q)
            assert len(ends) == 2, "2-D faces of a prism project to a segment"
            self.face_ends_2d.append((ends[0], ends[1]))
        # polygon: the face looking along +y (assembly_env.py:100-112)
        self.polygon_2d = [tuple(v) for v in self._vertices_2d_from_mesh()]
        (x0, y0, z0), (x1, y1, z1) = self.bounding_box
        self.depth = y1 - y0
        # area and area centroid (shoelace) -- mass properties for equilibrium
        a2 = cx = cz = 0.0
        poly = self.polygon_2d
        for i in range(len(poly)):
            xa, za = poly[i]
            xb, zb = poly[(i + 1) % len(poly)]
            w = xa * zb - xb * za
            a2 += w
            cx += (xa + xb) * w
            cz += (za + zb) * w
        self.area = abs(a2) / 2.0
        self.centroid_2d = (cx / (3.0 * a2), cz / (3.0 * a2))
        # largest centre-of-mass -> vertex distance: length unit of the torque rows
        self.radius = max(math.sqrt((x - self.centroid_2d[0]) ** 2 + (z - self.centroid_2d[1]) ** 2) for x, z in poly)

    def _vertices_2d_from_mesh(self):
        vertices = None
        for face, vertices in self.mesh.face.items():
            n = self.mesh.face_normal(face)
            if abs(n[1] - 1) < 1e-3:
                break
        for i in vertices:
            v = self.mesh.vertex_coordinates(i)
            yield [v[0], v[2]]

    # -- reference API
    @property
    def num_faces(self):
        return len(self._all_faces)

    @property
    def faces(self):
        return self._all_faces

    @property
    def faces_2d(self):
        return range(self.num_faces_2d)

    @property
    def target_faces_2d(self):
        return self._target_faces_2d or self.faces_2d

    @property
    def receiving_faces_2d(self):
        return self._receiving_faces_2d or self.faces_2d

    @property
    def num_faces_2d(self):
        return len(self._2d_faces)

    @property
    def vertices(self):
        for key in self.mesh.vertices():
            yield self.mesh.vertex_coordinates(key)

    @property
    def vertices_2d(self):
        for v in self.polygon_2d:
            yield [v[0], v[1]]

    def get_face_frame_2d(self, face):
        """(point (x, z), normal (x, z)) of assembly_env.py:118-124; the frame's
        x axis is (n_z, -n_x) and its z axis the face normal."""
        return self.face_centers_2d[face], self.face_normals_2d[face]

    def contains_2d(self, points):
        """assembly_env.py:126-137 on an (N, 2) float64 array; the half-plane value
        is (px-cx)*nx + (pz-cz)*nz with separately rounded operations."""
        import numpy as np
        contains = np.ones(len(points), dtype=bool)
        px = points[:, 0]
        pz = points[:, 1]
        for (cx, cz), (nx, nz) in zip(self.face_centers_2d, self.face_normals_2d):
            value = (px - cx) * nx + (pz - cz) * nz
            contains = contains & (value <= 0)
        return contains


class Block(Shape):
    """assembly_env.py:140-156 in canonical 2-D form."""

    def __init__(self, shape, position, orientation=None, object_id=None, pose=None):
        self.shape = shape
        self.object_id = object_id
        self.is_static = False
        if pose is None:
            if orientation is None:
                c, s = 1.0, 0.0
            else:
                # quaternion (w, x, y, z) of a rotation about y: matrix entries
                # m00 = 1 - 2(y^2 + z^2), m02 = 2(xz + wy)
                w, x, y, z = orientation
                c = 1.0 - 2.0 * (y * y + z * z)
                s = 2.0 * (x * z + w * y)
            pose = (float(position[0]), float(position[2]), c, s)
        self.pose = tuple(float(v) for v in pose)
        tx, tz, c, s = self.pose
        self.position = [tx, 0.0 if position is None else float(position[1]), tz]
        self.orientation = orientation
        # Shape.__init__(mesh=posed mesh) of the reference, restated on the tables
        self.urdf_file = shape.urdf_file
        self.name = shape.name
        self._target_faces_2d = None          # not inherited (assembly_env.py:153)
        self._receiving_faces_2d = None
        self._2d_faces = shape._2d_faces
        self._all_faces = shape._all_faces
        self.depth = shape.depth
        self.area = shape.area
        self.radius = shape.radius
        self.face_normals_2d = [rot(c, s, nx, nz) for nx, nz in shape.face_normals_2d]
        self.face_centers_2d = [self._apply(p) for p in shape.face_centers_2d]
        self.face_ends_2d = [(self._apply(a), self._apply(b)) for a, b in shape.face_ends_2d]
        self.polygon_2d = [self._apply(p) for p in shape.polygon_2d]
        self.centroid_2d = self._apply(shape.centroid_2d)
        xs = [p[0] for p in self.polygon_2d]
        zs = [p[1] for p in self.polygon_2d]
        (_, y0, _), (_, y1, _) = shape.bounding_box
        self.bounding_box = ((min(xs), y0, min(zs)), (max(xs), y1, max(zs)))

    def _apply(self, p):
        tx, tz, c, s = self.pose
        x, z = rot(c, s, p[0], p[1])
        return (x + tx, z + tz)

    @property
    def vertices(self):
        (_, y0, _), (_, y1, _) = self.bounding_box
        for y in (y0, y1):
            for x, z in self.polygon_2d:
                yield [x, y, z]

    def __repr__(self):
        return f"Block ({self.object_id})"


def polygon_separation(a, b):
    """Signed separation of two posed convex polygons (negative = penetration depth): the largest,
    over the faces of both, of the smallest half-plane value of the other polygon's vertices."""
    sep = -math.inf
    for p, q in ((a, b), (b, a)):
        for (cx, cz), (nx, nz) in zip(p.face_centers_2d, p.face_normals_2d):
            s = min((vx - cx) * nx + (vz - cz) * nz for vx, vz in q.polygon_2d)
            if s > sep:
                sep = s
    return sep


class AssemblyEnv:
    """assembly_env.py:159-438, default back-ends (CRA model, no PyBullet)."""

    def __init__(self, render=False, bounds=None, stability="rbe", mu=0.8, density=1.0, cra_env=True,
                 pybullet_env=False, tmax=1e-6, amin=1e-3, collision_tol=0.005):
        from . import stability as st
        # pybullet_env=True: the collision flags of `_check_collision` (assembly_env.py:346-391) are
        # produced -- by exact convex-polygon penetration depths instead of Bullet's contact points
        # (PARITY UNPINNED against Bullet: pybullet==3.2.6 is not in /root/reference nor installed).
        # The PyBullet settling check `is_stable_pybullet` stays out of scope.
        self.collision_enabled = bool(pybullet_env)
        self.collision_tol = collision_tol
        if bounds is None:
            bounds = [[-3.0, -3.0, -1.0], [7.0, 7.0, 9.0]]
        self.bounds = bounds
        self.obstacles = []
        self.blocks = []
        self._state_info = None
        self.mu = mu
        self.density = density
        self.tmax = tmax
        self.amin = amin
        self.client = None
        if stability == "rbe":
            self.stability_fct = st.is_stable_rbe
        elif stability is None:
            self.stability_fct = lambda env: (None, None)
        elif stability == "pybullet":
            raise NotImplementedError("PyBullet back-end is out of scope")
        else:
            self.stability_fct = stability
        self.cra_env = cra_env
        self.cra_assembly = None
        self.num_interface_extractions = 0
        self.num_stability_solves = 0
        self.reset()

    def reset(self):
        self.obstacles = []
        self.blocks = []
        self.is_block_frozen = False
        self.frozen_block_index = None
        self._reset_cra_assembly()
        self._update_state_info()

    def _reset_cra_assembly(self):
        """assembly_env.py:281-304: rebuild the assembly and detect interfaces."""
        from . import stability as st
        if not self.cra_env:
            return
        self.cra_assembly = st.CRAAssembly(self.bounds, self.blocks, tmax=self.tmax, amin=self.amin)
        self.num_interface_extractions += 1

    def _check_collision(self):
        """assembly_env.py:346-391 for the last block: bounds test on the block POSITION (:360),
        then penetration deeper than `tol` against every other block, the floor and the obstacles.
        Bullet's `contact distance < -tol` is restated as the exact penetration depth of two convex
        polygons (separating-axis form): sep = max over the faces of both polygons of
        min over the other polygon's vertices of (v - c).n, collision iff sep < -tol; the floor is
        the half-plane z <= 0.  Canonical arithmetic: (vx-cx)*nx + (vz-cz)*nz, separately rounded."""
        info = {"obstacles": [], "blocks": [], "floor": False, "bounding_box": False}
        if len(self.blocks) == 0:
            return False, info
        block = self.blocks[-1]
        tol = self.collision_tol
        lo, hi = self.bounds
        if any(block.position[k] < lo[k] for k in range(3)) or any(block.position[k] > hi[k] for k in range(3)):
            info["bounding_box"] = True
        for i, b in enumerate(self.blocks[:-1]):
            if polygon_separation(b, block) < -tol:
                info["blocks"].append(i)
        if min(v[1] for v in block.polygon_2d) < -tol:
            info["floor"] = True
        for i, obs in enumerate(self.obstacles):
            if polygon_separation(obs, block) < -tol:
                info["obstacles"].append(i)
        return any(bool(v) for v in info.values()), info

    def _update_state_info(self):
        if self.collision_enabled:
            collision, collision_info = self._check_collision()
        else:                                   # assembly_env.py:310-312
            collision = False
            collision_info = {"obstacles": [], "blocks": [], "floor": False, "bounding_box": False}
        self._state_info = {
            "last_block": self.blocks[-1] if self.blocks else None,
            "collision": collision,
            "collision_info": collision_info,
            "frozen_block": self.frozen_block_index,
        }
        is_stable, stability_info = self.stability_fct(self)
        self.num_stability_solves += 1
        self._state_info["stable"] = is_stable
        self._state_info["stability_info"] = stability_info

    def add_block(self, block):
        self.blocks.append(block)
        self._reset_cra_assembly()
        self._update_state_info()
        return self._state_info

    @property
    def state_info(self):
        return self._state_info

    def get_floor_frame(self):
        """Frame.worldXY(): point (0, 0), normal +z."""
        return (0.0, 0.0), (0.0, 1.0)

    def add_obstacle(self, obstacle):
        self.obstacles.append(obstacle)

    def is_stable(self):
        return self._state_info["stable"]

    def freeze_block(self, block_index):
        self.blocks[block_index].is_static = True
        if self.cra_assembly is not None:
            self.cra_assembly.set_boundary_condition(block_index)

    def unfreeze_block(self, block_index):
        self.blocks[block_index].is_static = False
        if self.cra_assembly is not None:
            self._reset_cra_assembly()
