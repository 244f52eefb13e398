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


def render_blocks_2d(blocks, xlim, ylim, img_size=(64, 64)):
    if tuple(img_size) != (L.BW_IMG, L.BW_IMG):
        raise NotImplementedError("bridges_b200 renders 64 x 64 rasters (successor_dqn.py --img_size default)")
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
