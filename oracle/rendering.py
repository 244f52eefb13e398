"""Oracle restatement of `render_blocks_2d` (assembly_gym/utils/rendering.py:105-113).

TEST INFRASTRUCTURE (see oracle/__init__.py).
"""
import numpy as np


def pixel_grid(xlim, ylim, img_size):
    X, Y = np.meshgrid(np.linspace(*xlim, img_size[0]), np.linspace(ylim[1], ylim[0], img_size[1]))
    return np.vstack([X.ravel(), Y.ravel()]).T


def render_blocks_2d(blocks, xlim, ylim, img_size=(512, 512)):
    image = np.zeros(img_size, dtype=bool)
    # Y axis reversed: row 0 is the top of the scene
    positions = pixel_grid(xlim, ylim, img_size)
    for block in blocks:
        image = image | block.contains_2d(positions).reshape(img_size)
    return image
