"""Drop-in `Shape`, `Block`, `AssemblyEnv` (assembly_gym/assembly_gym/envs/assembly_env.py).

These are thin host-side views: every number that decides an observation (poses, rasters,
bounds flags, interfaces, verdicts, rewards) is produced by the CUDA library; the classes
only hold what the reference's callers read (`.blocks`, `.position`, `.vertices_2d`,
`.get_face_frame_2d`, `.state_info`, `.mu`, `.bounds`, ...).
"""
import numpy as np

from ..shapes_io import load_shape_tables


class Frame2D:
    """The subset of compas' Frame the callers of `get_face_frame_2d` use."""

    def __init__(self, point, normal):
        self.point = [point[0], 0.0, point[1]]
        self.normal = [normal[0], 0.0, normal[1]]
        self.xaxis = [normal[1], 0.0, -normal[0]]
        self.yaxis = [0.0, 1.0, 0.0]
        self.zaxis = self.normal

    def to_world_coordinates(self, local):
        x, y, z = local
        return [self.point[0] + x * self.xaxis[0] + z * self.normal[0], self.point[1] + y,
                self.point[2] + x * self.xaxis[2] + z * self.normal[2]]


class Shape:
    """assembly_env.py:21-137 (loading by shapes_io; `mesh=` construction is not supported)."""

    def __init__(self, mesh=None, urdf_file=None, name="", receiving_faces_2d=None, target_faces_2d=None, tables=None):
        if mesh is not None:
            raise NotImplementedError("bridges_b200 shapes are loaded from URDF files")
        self.name = name
        self.urdf_file = None
        self.tables = tables
        if urdf_file is not None:
            self.tables = load_shape_tables(urdf_file)
            self.urdf_file = self.tables.urdf_file
        self._target_faces_2d = target_faces_2d
        self._receiving_faces_2d = receiving_faces_2d
        if self.tables is not None:
            self.bounding_box = self.tables.aabb

    @property
    def num_faces_2d(self):
        return len(self.tables.normals)

    @property
    def num_faces(self):
        return len(self.tables.face_keys)

    @property
    def faces(self):
        return self.tables.face_keys

    @property
    def faces_2d(self):
        return range(self.num_faces_2d)

    @property
    def target_faces_2d(self):
        return self._target_faces_2d or self.faces_2d

    @property
    def receiving_faces_2d(self):
        return self._receiving_faces_2d or self.faces_2d

    def _xz(self, p):
        return p

    @property
    def vertices(self):
        for v in self.tables.vertices3d:
            x, z = self._xz((v[0], v[2]))
            yield [x, v[1], z]

    @property
    def vertices_2d(self):
        for p in self.tables.polygon:
            x, z = self._xz(p)
            yield [x, z]

    def contains_2d(self, points):
        """assembly_env.py:126-137, evaluated by the CUDA library (same arithmetic as the rasters)."""
        from ..utils.rendering import contains_2d
        return contains_2d(self, points)

    def get_face_frame_2d(self, face):
        c = self._xz(self.tables.centers[face])
        n = self._dir(self.tables.normals[face])
        return Frame2D(c, n)

    def _dir(self, n):
        return n


class Block(Shape):
    """assembly_env.py:140-156.  `pose` = (x, z, cos, sin) as returned by the CUDA placement."""

    def __init__(self, shape, position, orientation=None, object_id=None, pose=None):
        super().__init__(name=shape.name, tables=shape.tables)
        self.shape = shape
        self.urdf_file = shape.urdf_file
        self.object_id = object_id
        self.is_static = False
        if pose is None:
            c, s = 1.0, 0.0
            if orientation is not None:
                w, x, y, z = (orientation if not hasattr(orientation, "wxyz") else orientation.wxyz)
                c = 1.0 - 2.0 * (y * y + z * z)
                s = 2.0 * (x * z + w * y)
            pose = (float(position[0]), float(position[2]), c, s)
        self.pose = tuple(float(v) for v in pose)
        self.position = [self.pose[0], float(position[1]) if position is not None else 0.0, self.pose[1]]
        self.orientation = orientation if orientation is not None else self.quaternion
        xs, zs = zip(*[self._xz(p) for p in self.tables.polygon])
        self.bounding_box = ((min(xs), self.tables.ymin, min(zs)), (max(xs), self.tables.ymax, max(zs)))

    @property
    def quaternion(self):
        """(w, x, y, z) of the rotation about y with cos = c, sin = s."""
        _, _, c, s = self.pose
        w = float(np.sqrt(max(0.0, (1.0 + c) / 2.0)))
        y = float(np.sqrt(max(0.0, (1.0 - c) / 2.0))) * (1.0 if s >= 0 else -1.0)
        return (w, 0.0, y, 0.0)

    def _xz(self, p):
        tx, tz, c, s = self.pose
        return (c * p[0] + s * p[1] + tx, c * p[1] - s * p[0] + tz)

    def _dir(self, n):
        _, _, c, s = self.pose
        return (c * n[0] + s * n[1], c * n[1] - s * n[0])

    def __repr__(self):
        return f"Block ({self.object_id})"


class AssemblyEnv:
    """assembly_env.py:159-438: configuration holder + world-state view.  The state itself lives
    on the GPU inside the `AssemblyGym` that owns this object."""

    def __init__(self, render=False, bounds=None, stability="rbe", mu=0.8, density=1.0, cra_env=True,
                 pybullet_env=False, device=0):
        if stability == "pybullet":
            raise NotImplementedError("the PyBullet settling check is not part of bridges_b200 (SURVEY.md section 8f)")
        # pybullet_env=True switches the collision flags of `_check_collision` (assembly_env.py:346-391) on;
        # they come from the CUDA step (exact polygon penetration depths, tol 0.005), not from Bullet
        self.collision = bool(pybullet_env)
        self.collision_tol = 0.005
        if stability not in ("rbe", None):
            raise NotImplementedError("stability must be 'rbe' (default) or None")
        if bounds is None:
            bounds = np.array([[-3., -3., -1], [7., 7., 9.]])
        self.bounds = np.asarray(bounds, dtype=float)
        self.mu = mu
        self.density = density
        self.stability = stability
        self.device = device
        self.client = None
        self.cra_assembly = None
        self.obstacles = []
        self.blocks = []
        self.is_block_frozen = False
        self.frozen_block_index = None
        self._gym = None
        self._state_info = {"last_block": None, "collision": False,
                            "collision_info": {"obstacles": [], "blocks": [], "floor": False, "bounding_box": False},
                            "frozen_block": None, "stable": True if stability else None, "stability_info": None}

    @property
    def state_info(self):
        return self._state_info

    def is_stable(self):
        return self._state_info["stable"]

    def get_floor_frame(self):
        return Frame2D((0.0, 0.0), (0.0, 1.0))

    def disconnect_client(self):
        pass

    # the mutators below are routed through the owning AssemblyGym (GPU state)
    def reset(self):
        if self._gym is not None:
            self._gym._reset_world()

    def add_block(self, block):
        return self._require_gym()._add_block(block)

    def add_obstacle(self, obstacle):
        self.obstacles.append(obstacle)

    def freeze_block(self, block_index):
        self._require_gym()._set_static(block_index, True)

    def unfreeze_block(self, block_index):
        self._require_gym()._set_static(block_index, False)

    def _update_state_info(self):
        self._require_gym()._evaluate()

    def _require_gym(self):
        if self._gym is None:
            raise RuntimeError("this AssemblyEnv is not attached to an AssemblyGym yet")
        return self._gym
