"""Drop-in `render_blocks_2d` (assembly_gym/assembly_gym/utils/rendering.py:105-113), rastered
by the CUDA library (64 x 64 only, the size the training script uses)."""
import ctypes as C

import numpy as np

from .. import lib as L
from ..envs.batched import BatchedAssemblyGym, shape_desc

_renderer = None


def _get_renderer(device=0):
    global _renderer
    if _renderer is None:
        _renderer = BatchedAssemblyGym(1, ["shapes/cube06.urdf"], device=device)
    return _renderer


def contains_2d(shape_or_block, points):
    """Shape.contains_2d / Block.contains_2d (assembly_env.py:126-137): points [n, 2] (x, z) -> bool [n]."""
    core = _get_renderer()
    pts = np.ascontiguousarray(np.asarray(points, dtype=np.float64).reshape(-1, 2))
    desc = shape_desc(shape_or_block.tables)
    pose = getattr(shape_or_block, "pose", None)
    blk = None
    if pose is not None:
        blk = np.zeros(1, dtype=core.dt["block"])
        blk[0] = (pose[0], pose[1], pose[2], pose[3], 0, 0)
    out = np.zeros(len(pts), dtype=np.uint8)
    core._check(core.lib.bw_contains_2d_host(core.handle, C.byref(desc), blk.ctypes.data if blk is not None else None,
                                             pts.ctypes.data, len(pts), out.ctypes.data))
    return out.astype(bool)


def render_blocks_2d(blocks, xlim, ylim, img_size=(64, 64)):
    """rendering.py:105-113.  64 x 64 (the training script's size) takes the raster kernel; any other size is the
    reference's own construction -- contains_2d of every block on the meshgrid of pixel nodes."""
    if tuple(img_size) != (L.BW_IMG, L.BW_IMG):
        X, Y = np.meshgrid(np.linspace(xlim[0], xlim[1], img_size[1]), np.linspace(ylim[1], ylim[0], img_size[0]))
        pts = np.array([X.flatten(), Y.flatten()]).T
        img = np.zeros(len(pts), dtype=bool)
        for b in blocks:
            img |= contains_2d(b, pts)
        return img.reshape(img_size)
    core = _get_renderer()
    blocks = list(blocks)
    tables, index = [], {}
    for b in blocks:
        key = id(b.tables)
        if key not in index:
            index[key] = len(tables)
            tables.append(b.tables)
    if len(tables) > L.BW_MAX_SHAPES:
        raise L.BridgesError("too many distinct shapes in one render call")
    descs = (L.bw_shape_desc * max(1, len(tables)))()
    for i, t in enumerate(tables):
        descs[i] = shape_desc(t)
    arr = np.zeros(max(1, len(blocks)), dtype=core.dt["block"])
    for i, b in enumerate(blocks):
        arr[i] = (b.pose[0], b.pose[1], b.pose[2], b.pose[3], index[id(b.tables)], 0)
    bits = np.zeros(L.BW_IMG, dtype=np.uint64)
    xl = np.asarray(xlim, dtype=np.float64)
    yl = np.asarray(ylim, dtype=np.float64)
    core._check(core.lib.bw_render_blocks_host(core.handle, descs, len(tables), arr.ctypes.data, len(blocks),
                                               xl.ctypes.data, yl.ctypes.data, bits.ctypes.data))
    return BatchedAssemblyGym.bits_to_bool(bits)


def _plotting(*args, **kwargs):
    raise NotImplementedError("matplotlib / PyBullet plotting helpers are debug visualisation, not part of bridges_b200 "
                              "(SURVEY.md section 2, row 5)")


# names the reference's training script imports next to render_blocks_2d (successor_dqn.py:14)
get_rgb_array = plot_cra_assembly = render_assembly_env = plot_assembly_env = _plotting
