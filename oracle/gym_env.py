"""Oracle restatement of assembly_gym/assembly_gym/envs/gym_env.py.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Same class / function names as
the reference so that the parity tests read like the reference's own scripts
(utils/structures.py, utils/test_suite.py, the notebooks).
"""
from dataclasses import dataclass

import numpy as np

from .assembly_env import AssemblyEnv, Block, Shape
from .geometry import align_frames_2d, box_contains_point, distance_box_point


def sparse_reward(gym_env, obs, info):
    """gym_env.py:11-22."""
    if gym_env.assembly_env.state_info["collision"] or not gym_env.assembly_env.state_info["stable"]:
        return -1
    num_targets_reached = len(obs["targets_reached"])
    if not gym_env.all_targets_reached():
        return -1 + num_targets_reached
    return num_targets_reached


def _library(trapezoid, hexagon):
    shapes = []
    if trapezoid:
        shapes.append(Shape(urdf_file="shapes/trapezoid.urdf", name="trapezoid"))
    if hexagon:
        shapes.append(Shape(urdf_file="shapes/hexagon.urdf", name="hexagon"))
    return shapes


def horizontal_bridge_setup(square_size=0.6, num_obstacles=5, trapezoid=True, hexagon=False):
    """gym_env.py:25-43."""
    shapes = _library(trapezoid, hexagon)
    reward_x = num_obstacles * square_size + 2.5 * square_size
    targets = [(reward_x, 0, square_size / 2)]
    obstacles = [(i * square_size, 0, square_size / 2) for i in range(1, num_obstacles + 1)]
    return dict(shapes=shapes, obstacles=obstacles, targets=targets)


def bridge_setup(H=0.8, num_stories=1, trapezoid=True, hexagon=False):
    """gym_env.py:46-61."""
    shapes = _library(trapezoid, hexagon)
    targets = [(0.5, 0, num_stories * H + H / 2)]
    obstacles = [(targets[0][0], 0.0, i * H + H / 2) for i in range(num_stories)]
    return dict(shapes=shapes, obstacles=obstacles, targets=targets)


def tower_setup(num_targets=3, targets=None, rng=None):
    """gym_env.py:64-79 (random targets drawn from `rng` instead of the global numpy state)."""
    if targets is None:
        rng = rng or np.random
        targets = [(rng.uniform(-4, 4), 0, rng.uniform(0.0, 4)) for _ in range(num_targets)]
    shapes = [Shape(urdf_file="shapes/trapezoid.urdf", name="trapezoid")]
    return dict(shapes=shapes, obstacles=[], targets=targets)


def hard_tower_setup():
    """gym_env.py:82-88."""
    trapezoid = Shape(urdf_file="shapes/trapezoid.urdf", name="trapezoid")
    cube = Shape(urdf_file="shapes/cube1.urdf", name="cube", receiving_faces_2d=[0], target_faces_2d=[2])
    return dict(shapes=[trapezoid, cube], targets=[[0, 0, 0.5], [0, 0, 5.5]], obstacles=[[0, 0, 2.0]])


def tower_height_setup(tower_height=2, square_size=0.6):
    """The `--tower_height=k` task of BASELINE.json (the flag is absent from the
    reference snapshot; definition from SURVEY.md section 8(d).3): one column of
    k-1 obstacle cubes at x = 0.6 and one target just above it."""
    shapes = [Shape(urdf_file="shapes/trapezoid.urdf", name="trapezoid")]
    obstacles = [(square_size, 0, i * square_size + square_size / 2) for i in range(tower_height - 1)]
    targets = [(square_size, 0, (tower_height - 1) * square_size + square_size / 2)]
    return dict(shapes=shapes, obstacles=obstacles, targets=targets)


@dataclass
class Action:
    """gym_env.py:102-110."""
    target_block: int
    target_face: int
    shape: int
    face: int
    offset_x: float = 0.
    offset_y: float = 0.
    frozen: bool = False


class AssemblyGym:
    """gym_env.py:112-333."""

    def __init__(self, reward_fct, shapes=None, obstacles=None, targets=None, render_mode=None,
                 assembly_env=None, restrict_2d=False, max_steps=None):
        self.blocks = []
        self.shapes = []
        self.obstacles = []
        self.targets = []
        self.reward_fct = reward_fct
        self.render_mode = render_mode
        self.restrict_2d = restrict_2d
        self.action_history = None
        self.block_graph = None
        self.max_steps = max_steps
        if not restrict_2d:
            raise NotImplementedError
        if assembly_env is None:
            assembly_env = AssemblyEnv(render=render_mode == "human")
        self.assembly_env = assembly_env
        self.reset(shapes, obstacles, targets)

    def terminated(self, assembly_env):
        terminated = (not assembly_env.state_info["stable"] or assembly_env.state_info["collision"]
                      or self.all_targets_reached())
        truncated = self.max_steps and len(self.blocks) >= self.max_steps
        return terminated, truncated

    @property
    def num_targets(self):
        return len(self.targets)

    @property
    def num_obstacles(self):
        return len(self.obstacles)

    def distance_to_targets(self):
        if len(self.assembly_env.blocks) == 0:
            return self.num_targets * [np.inf]
        return [min(distance_box_point(block.bounding_box, target) for block in self.assembly_env.blocks)
                for target in self.targets]

    def _update_targets(self, new_block):
        # removes while iterating, exactly as gym_env.py:162-168 (quirk 8 of SURVEY App. C)
        targets_reached = []
        for target in self.targets_remaining:
            if box_contains_point(new_block.bounding_box, target):
                self.targets_reached.append(target)
                self.targets_remaining.remove(target)
        return targets_reached

    def all_targets_reached(self):
        return len(self.targets_remaining) == 0

    def _get_obs(self):
        info = self.assembly_env.state_info
        return {
            "blocks": self.blocks,
            "stable": bool(info["stable"]),
            "collision": bool(info["collision"]),
            "collision_block": bool(info["collision_info"]["blocks"]),
            "collision_obstacle": bool(info["collision_info"]["obstacles"]),
            "collision_floor": bool(info["collision_info"]["floor"]),
            "collision_boundary": bool(info["collision_info"]["bounding_box"]),
            "frozen_block": self.assembly_env.frozen_block_index,
            "obstacles": self.obstacles,
            "obstacle_blocks": self.assembly_env.obstacles,
            "targets": self.targets,
            "targets_remaining": self.targets_remaining,
            "targets_reached": self.targets_reached,
            "distance_to_targets": self.distance_to_targets(),
        }

    def _get_info(self):
        return {"blocks_initial_state": None, "blocks_final_state": None}

    def create_block(self, action):
        """gym_env.py:204-216."""
        if action.target_block == -1:
            block_frame = self.assembly_env.get_floor_frame()
        else:
            block_frame = self.assembly_env.blocks[action.target_block].get_face_frame_2d(action.target_face)
        shape_frame = self.shapes[action.shape].get_face_frame_2d(action.face)
        offset = [action.offset_x, 0, action.offset_y]
        (tx, tz), (c, s) = align_frames_2d(block_frame, shape_frame, offset)
        return Block(self.shapes[action.shape], position=[tx, 0.0, tz], pose=(tx, tz, c, s))

    def step(self, action):
        """gym_env.py:218-253."""
        new_block = self.create_block(action)
        self.assembly_env.add_block(new_block)
        self.action_history.append(action)
        self.blocks.append(new_block)
        new_block_index = len(self.assembly_env.blocks) - 1
        key = (action.target_block, action.target_face)
        if key not in self.block_graph:
            self.block_graph[key] = []
        self.block_graph[key].append((new_block_index, action.face))
        self.block_graph[(new_block_index, action.face)] = [key]
        if len(self.assembly_env.blocks) > 1 and self.assembly_env.blocks[-2].is_static:
            self.assembly_env.unfreeze_block(len(self.assembly_env.blocks) - 2)
        action.frozen = True
        if action.frozen:
            self.assembly_env.freeze_block(len(self.assembly_env.blocks) - 1)
        self._update_targets(new_block)
        self.assembly_env._update_state_info()
        terminated, truncated = self.terminated(self.assembly_env)
        info = self._get_info()
        observation = self._get_obs()
        reward = self.reward_fct(self, observation, info)
        return observation, reward, terminated, truncated, info

    def reset(self, shapes=None, obstacles=None, targets=None, blocks=None):
        """gym_env.py:255-289."""
        self.assembly_env.reset()
        self.action_history = []
        self.blocks = []
        self.block_graph = {(-1, 0): []}
        self.targets_reached = []
        if shapes is not None:
            self.shapes = shapes
        if obstacles is not None:
            self.obstacles = obstacles
        if targets is not None:
            self.targets = targets
        if blocks is not None:
            self.blocks = blocks
        self.targets_remaining = list(self.targets).copy()
        small_cube = Shape(urdf_file="shapes/cube06.urdf")
        for b in self.blocks:
            self.assembly_env.add_block(Block(self.shapes[b[-1]], b[:3], tuple(b[3:7])))
        for position in self.obstacles:
            self.assembly_env.add_obstacle(Block(shape=small_cube, position=position))
        return self._get_obs(), self._get_info()

    @property
    def num_step(self):
        return len(self.action_history)

    def collision_on_action(self, action, xlim, ylim):
        """gym_env.py:304-323."""
        block = self.create_block(action)
        eps = 1e-6
        collisions = False
        for vertex in block.vertices:
            if (vertex[0] < xlim[0] - eps or vertex[0] > xlim[1] + eps or
                    vertex[2] < ylim[0] - eps or vertex[2] > ylim[1] + eps):
                collisions = True
                break
        for vertex in block.vertices:
            if vertex[2] < -eps:
                collisions = True
                break
        return collisions

    def stabilities_freezing(self):
        """gym_env.py:325-333."""
        self.assembly_env._update_state_info()
        stable = self.assembly_env.is_stable()
        self.assembly_env.unfreeze_block(len(self.assembly_env.blocks) - 1)
        self.assembly_env._update_state_info()
        unfreezestable = self.assembly_env.is_stable()
        self.assembly_env.freeze_block(len(self.assembly_env.blocks) - 1)
        self.assembly_env._update_state_info()
        return stable, unfreezestable
