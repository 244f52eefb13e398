"""Oracle restatement of robotoddler/utils/actions.py (generate_actions :7-52,
filter_actions :71-82) with numpy arrays in place of torch tensors.

TEST INFRASTRUCTURE (see oracle/__init__.py).
"""
import numpy as np

from .gym_env import Action


def generate_actions(gym, x_discr_ground, offset_values=None, max_angle_rad=2 * np.pi + 0.1, max_blocks_per_face=1):
    if offset_values is None:
        offset_values = [0.]
    for shape_index in range(len(gym.shapes)):
        shape = gym.shapes[shape_index]
        for face in shape.target_faces_2d:
            for offset_x in x_discr_ground:
                yield Action(-1, 0, shape_index, face, offset_x, offset_y=0.)
            for target_block in range(len(gym.assembly_env.blocks)):
                block = gym.assembly_env.blocks[target_block]
                for target_face in block.receiving_faces_2d:
                    _, normal = block.get_face_frame_2d(target_face)
                    angle = np.arccos(min(1.0, max(-1.0, normal[1])))
                    if max_angle_rad is not None and angle > max_angle_rad:
                        continue
                    if max_blocks_per_face and len(gym.block_graph.get((target_block, target_face), tuple())) >= max_blocks_per_face:
                        continue
                    for offset_x in offset_values:
                        yield Action(target_block, target_face, shape_index, face, offset_x, offset_y=0.)


def filter_actions(gym_env, available_actions, action_features, block_features, obstacle_features, xlim, ylim):
    mask = np.zeros(len(available_actions), dtype=bool)
    reduced = []
    for i, action in enumerate(available_actions):
        if (not gym_env.collision_on_action(action, xlim, ylim)
                and np.sum(action_features[i] * block_features) == 0
                and np.sum(action_features[i] * obstacle_features) == 0):
            mask[i] = True
            reduced.append(action)
    return reduced, action_features[mask], mask
