"""Oracle restatement of assembly_gym/assembly_gym/utils/geometry.py (hot-path part).

TEST INFRASTRUCTURE (see oracle/__init__.py).
"""
import math


def align_frames_2d(frame1, frame2, frame1_coordinates=None):
    """geometry.py:39-50 in canonical 2-D arithmetic.

    frame = ((px, pz), (nx, nz)).  The reference rotates about
    `cross(n1, n2) + (0, 1e-6, 0)` by `arccos(clip(-n1.n2, -1, 1))`; both normals
    lie in the xz-plane, so the axis is +-y and the rotation is, about +y,

        c = clip(-(n1.n2)),   s = sign(cross_y + 1e-6) * |cross_y|,
        cross_y = n1z*n2x - n1x*n2z,

    i.e. the closed form of cos/sin(arccos(.)) without libm calls (which the GPU
    could not reproduce bit for bit).  The 1e-6 axis hack is kept: for
    cross_y in (-1e-6, 0) the rotation sense flips exactly as in the reference.
    Returns (offset (x, z), (c, s)).
    """
    if frame1_coordinates is None:
        frame1_coordinates = [0, 0, 0]
    (p1x, p1z), (n1x, n1z) = frame1
    (p2x, p2z), (n2x, n2z) = frame2
    ox, _, oy = frame1_coordinates
    c = -(n1x * n2x + n1z * n2z)
    c = min(1.0, max(-1.0, c))
    cross_y = n1z * n2x - n1x * n2z
    s = abs(cross_y)
    if not (cross_y + 1e-6 > 0.0):
        s = -s
    # frame1.to_world_coordinates([ox, 0, oy]): point + ox*xaxis + oy*zaxis, xaxis = (n1z, -n1x)
    wx = (n1z * ox + n1x * oy) + p1x
    wz = (n1z * oy - n1x * ox) + p1z
    # frame2.point.transformed(rotation)
    rx = c * p2x + s * p2z
    rz = c * p2z - s * p2x
    return (wx - rx, wz - rz), (c, s)


def box_contains_point(box, point, tol=1e-6):
    """compas Box.contains_point on an axis-aligned box ((xmin,ymin,zmin),(xmax,ymax,zmax))."""
    (x0, y0, z0), (x1, y1, z1) = box
    return (x0 - tol <= point[0] <= x1 + tol and
            y0 - tol <= point[1] <= y1 + tol and
            z0 - tol <= point[2] <= z1 + tol)


def project_point_on_box(box, point):
    """geometry.py:99-105."""
    (x0, y0, z0), (x1, y1, z1) = box
    return (min(max(point[0], x0), x1), min(max(point[1], y0), y1), min(max(point[2], z0), z1))


def distance_box_point(box, point):
    """geometry.py:89-96."""
    if box_contains_point(box, point):
        return 0.0
    q = project_point_on_box(box, point)
    dx, dy, dz = point[0] - q[0], point[1] - q[1], point[2] - q[2]
    return math.sqrt(dx * dx + dy * dy + dz * dz)
