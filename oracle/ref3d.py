"""Literal 3-D restatement of the reference's posing arithmetic -- TEST INFRASTRUCTURE.

The reference poses a block in 3-D through compas (`align_frames_2d` geometry.py:39-50 ->
`Rotation.from_axis_and_angle(...)` -> `.quaternion` -> `Rotation.from_quaternion` ->
`Translation * Rotation` -> `mesh.transformed` -> `Shape.from_mesh` of the posed mesh,
assembly_env.py:146-153) and uses libm `arccos/sin/cos`.  The canonical 2-D arithmetic of
oracle/assembly_env.py (and of the CUDA kernels) is the closed form of that chain.  This module
follows the chain step by step in float64 -- compas 2.1.1's published formulas for
`matrix_from_axis_and_angle`, `quaternion_from_matrix`, `matrix_from_quaternion`,
`transform_points`, `normal_polygon`, `centroid_points` -- so that tests can show the two agree to
rounding (poses to ~1e-15, rasters except for knife-edge pixels), i.e. that the canonical form
changes no decision the reference makes away from a measure-zero set.
"""
import math

import numpy as np

from . import compas_lite as cl


def matrix_from_axis_and_angle(axis, angle):
    """compas.geometry.matrix_from_axis_and_angle (3x3 part)."""
    axis = list(axis)
    if cl.length_vector(axis):
        axis = cl.normalize_vector(axis)
    sina = math.sin(angle)
    cosa = math.cos(angle)
    R = [[cosa, 0.0, 0.0], [0.0, cosa, 0.0], [0.0, 0.0, cosa]]
    outer = [[axis[i] * axis[j] * (1.0 - cosa) for i in range(3)] for j in range(3)]
    R = [[R[i][j] + outer[i][j] for i in range(3)] for j in range(3)]
    ax = [c * sina for c in axis]
    m = [[0.0, -ax[2], ax[1]], [ax[2], 0.0, -ax[0]], [-ax[1], ax[0], 0.0]]
    for i in range(3):
        for j in range(3):
            R[i][j] += m[i][j]
    return R


def quaternion_from_matrix(M):
    """compas.geometry.quaternion_from_matrix (w, x, y, z), Shepperd's branches."""
    qw2 = (1 + M[0][0] + M[1][1] + M[2][2]) / 4.0
    qx2 = (1 + M[0][0] - M[1][1] - M[2][2]) / 4.0
    qy2 = (1 - M[0][0] + M[1][1] - M[2][2]) / 4.0
    qz2 = (1 - M[0][0] - M[1][1] + M[2][2]) / 4.0
    qw2, qx2, qy2, qz2 = (max(v, 0.0) for v in (qw2, qx2, qy2, qz2))
    qw, qx, qy, qz = math.sqrt(qw2), math.sqrt(qx2), math.sqrt(qy2), math.sqrt(qz2)
    if qw >= qx and qw >= qy and qw >= qz:
        qx = (M[2][1] - M[1][2]) / (4 * qw)
        qy = (M[0][2] - M[2][0]) / (4 * qw)
        qz = (M[1][0] - M[0][1]) / (4 * qw)
    elif qx >= qw and qx >= qy and qx >= qz:
        qw = (M[2][1] - M[1][2]) / (4 * qx)
        qy = (M[0][1] + M[1][0]) / (4 * qx)
        qz = (M[0][2] + M[2][0]) / (4 * qx)
    elif qy >= qw and qy >= qx and qy >= qz:
        qw = (M[0][2] - M[2][0]) / (4 * qy)
        qx = (M[0][1] + M[1][0]) / (4 * qy)
        qz = (M[1][2] + M[2][1]) / (4 * qy)
    else:
        qw = (M[1][0] - M[0][1]) / (4 * qz)
        qx = (M[0][2] + M[2][0]) / (4 * qz)
        qy = (M[1][2] + M[2][1]) / (4 * qz)
    return [qw, qx, qy, qz]


def matrix_from_quaternion(q):
    """compas.geometry.matrix_from_quaternion (3x3 part)."""
    w, x, y, z = q
    n = w * w + x * x + y * y + z * z
    s = 2.0 / n
    xs, ys, zs = x * s, y * s, z * s
    wx, wy, wz = w * xs, w * ys, w * zs
    xx, xy, xz = x * xs, x * ys, x * zs
    yy, yz, zz = y * ys, y * zs, z * zs
    return [[1.0 - (yy + zz), xy - wz, xz + wy],
            [xy + wz, 1.0 - (xx + zz), yz - wx],
            [xz - wy, yz + wx, 1.0 - (xx + yy)]]


def apply(R, t, p):
    return [R[i][0] * p[0] + R[i][1] * p[1] + R[i][2] * p[2] + t[i] for i in range(3)]


class PosedMesh:
    """The posed mesh of a Block and the face frames the reference derives from it."""

    def __init__(self, shape, R, t):
        self.shape = shape
        self.R, self.t = R, t
        self.vertex = {k: apply(R, t, v) for k, v in shape.mesh.vertex.items()}
        self.faces_2d = shape._2d_faces

    def face_coordinates(self, fkey):
        return [self.vertex[k] for k in self.shape.mesh.face[fkey]]

    def face_frame_2d(self, face):
        """Shape.get_face_frame_2d (assembly_env.py:118-124): point = face centre, normal = face normal."""
        coords = self.face_coordinates(self.faces_2d[face])
        return cl.centroid_points(coords), cl.normal_polygon(coords)

    def contains_2d(self, points):
        """assembly_env.py:126-137 with numpy's dot, as the reference evaluates it."""
        contains = np.ones(len(points), dtype=bool)
        margin = np.full(len(points), np.inf)
        for k in range(len(self.faces_2d)):
            c, n = self.face_frame_2d(k)
            offset = np.array([c[0], c[2]])
            normal = np.array([n[0], n[2]])
            value = np.dot(points - offset, normal)
            contains = contains & (value <= 0)
            margin = np.minimum(margin, np.abs(value))
        return contains, margin


def align_frames_3d(frame1, frame2, frame1_coordinates):
    """geometry.py:39-50 verbatim: frames are (point, normal) 3-vectors; frame1's x axis is
    (n_z, 0, -n_x) (assembly_env.py:118-124; (1, 0, 0) for the floor frame)."""
    p1, n1 = frame1
    p2, n2 = frame2
    axis = cl.cross_vectors(n1, n2)
    axis = [axis[0], axis[1] + 1e-6, axis[2]]
    angle = math.acos(min(1.0, max(-1.0, -cl.dot_vectors(n1, n2))))
    R = matrix_from_axis_and_angle(axis, angle)
    xaxis = cl.normalize_vector([n1[2], 0.0, -n1[0]])
    ox, _, oy = frame1_coordinates
    world = [p1[i] + ox * xaxis[i] + oy * n1[i] for i in range(3)]
    rp2 = apply(R, [0.0, 0.0, 0.0], p2)
    offset = [world[i] - rp2[i] for i in range(3)]
    return offset, R


def place(shapes, blocks, action):
    """create_block (gym_env.py:204-216) + Block.__init__ (assembly_env.py:146-153) in 3-D.
    `blocks` is the list of PosedMesh placed so far; returns the new PosedMesh."""
    if action.target_block == -1:
        frame1 = ([0.0, 0.0, 0.0], [0.0, 0.0, 1.0])
    else:
        frame1 = blocks[action.target_block].face_frame_2d(action.target_face)
    shape = shapes[action.shape]
    mesh = shape.mesh
    coords = [mesh.vertex[k] for k in mesh.face[shape._2d_faces[action.face]]]
    frame2 = (cl.centroid_points(coords), cl.normal_polygon(coords))
    offset, R = align_frames_3d(frame1, frame2, [action.offset_x, 0.0, action.offset_y])
    # the reference stores rotation.quaternion and rebuilds the matrix from it
    R = matrix_from_quaternion(quaternion_from_matrix(R))
    return PosedMesh(shape, R, offset)
