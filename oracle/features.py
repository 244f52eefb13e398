"""Oracle restatement of the feature functions of robotoddler/training/successor_dqn.py
(get_state_features :47-64, get_task_features :67-85, get_action_features :88-94,
lin_reward :397-401) and robotoddler/utils/utils.py (gaussian_kernel / convolve_with_gaussian
:93-114).  Arrays are numpy; the Gaussian convolution uses torch.nn.functional.conv2d on the
CPU exactly as the reference does.

TEST INFRASTRUCTURE (see oracle/__init__.py).
"""
import numpy as np

from .assembly_env import Block, Shape
from .rendering import render_blocks_2d


def get_state_features(observation, xlim=(0, 1), ylim=(0, 1), img_size=(512, 512)):
    binary = np.array([observation['stable'], observation['collision'], observation['collision_block'],
                       observation['collision_obstacle'], observation['collision_floor'],
                       observation['collision_boundary']], dtype=np.float32)
    image = render_blocks_2d(observation['blocks'], xlim=xlim, ylim=ylim, img_size=img_size).astype(np.float32)
    return image[None], binary


def gaussian_kernel(kernel_size, sigma):
    import torch
    coords = torch.arange(kernel_size) - kernel_size // 2
    kernel1d = torch.exp(-(coords.float() ** 2) / (2 * sigma ** 2))
    kernel1d /= kernel1d.sum()
    return kernel1d.unsqueeze(0) * kernel1d.unsqueeze(1)


def convolve_with_gaussian(image, kernel_size, sigma):
    import torch
    import torch.nn.functional as F
    kernel = gaussian_kernel(kernel_size, sigma)
    out = F.conv2d(torch.from_numpy(image).unsqueeze(0).unsqueeze(0), kernel.unsqueeze(0).unsqueeze(0),
                   padding=kernel_size // 2)
    return out.squeeze(0).squeeze(0).numpy()


def get_task_features(obs, xlim=(0, 1), ylim=(0, 1), img_size=(512, 512)):
    cube = Shape(urdf_file='shapes/cube06.urdf')
    target_blocks = [Block(shape=cube, position=target) for target in obs['targets']]
    reward = render_blocks_2d(target_blocks, xlim=xlim, ylim=ylim, img_size=img_size).astype(np.float32)
    reward = convolve_with_gaussian(reward, 101, 16)
    obstacle = render_blocks_2d(obs['obstacle_blocks'], xlim=xlim, ylim=ylim, img_size=img_size).astype(np.float32)
    return reward[None], obstacle[None]


def get_action_features(env, actions, xlim=(0, 1), ylim=(0, 1), img_size=(512, 512)):
    blocks = [env.create_block(action) for action in actions]
    if not blocks:
        return np.zeros((0, 1) + tuple(img_size), dtype=np.float32)
    return np.array([render_blocks_2d([b], xlim=xlim, ylim=ylim, img_size=img_size) for b in blocks],
                    dtype=np.float32)[:, None]


def lin_reward(selected_action_features, reward_features, frozen_stable, unfrozen_stable):
    """successor_dqn.py:397-401 (float32 sum)."""
    value = np.float32(0.0)
    s = np.sum(selected_action_features * reward_features, dtype=np.float32)
    if frozen_stable:
        value = np.float32(s / np.float32(100))
    if unfrozen_stable:
        value = np.float32(s)
    return value
