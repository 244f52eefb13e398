"""Drop-in `AssemblyGym`, `Action`, `sparse_reward` and the task set-ups
(assembly_gym/assembly_gym/envs/gym_env.py) on top of the CUDA library, as a num_envs = 1
view of `BatchedAssemblyGym`.

    env = AssemblyGym(reward_fct=sparse_reward, max_steps=10, restrict_2d=True,
                      assembly_env=AssemblyEnv(render=False))
    obs, info = env.reset(**horizontal_bridge_setup(num_obstacles=3))
    obs, reward, terminated, truncated, info = env.step(Action(-1, 0, 0, 2, -0.45))
    frozen_stable, unfrozen_stable = env.stabilities_freezing()

is the reference's rollout code unchanged (successor_dqn.py:365-475).
"""
from dataclasses import dataclass

import numpy as np

from .. import lib as L
from .assembly_env import AssemblyEnv, Block, Shape
from .batched import BatchedAssemblyGym


def sparse_reward(gym_env, obs, info):
    """gym_env.py:11-22 (the value is computed by the step kernel; this function re-derives it
    from the observation the same way, so custom reward functions keep working)."""
    if gym_env.assembly_env.state_info['collision'] or not gym_env.assembly_env.state_info['stable']:
        return -1
    num_targets_reached = len(obs['targets_reached'])
    if not gym_env.all_targets_reached():
        return -1 + num_targets_reached
    return num_targets_reached


def _library(trapezoid, hexagon):
    shapes = []
    if trapezoid:
        shapes.append(Shape(urdf_file='shapes/trapezoid.urdf', name="trapezoid"))
    if hexagon:
        shapes.append(Shape(urdf_file='shapes/hexagon.urdf', name="hexagon"))
    return shapes


def horizontal_bridge_setup(square_size=0.6, num_obstacles=5, trapezoid=True, hexagon=False):
    """gym_env.py:25-43."""
    reward_x = num_obstacles * square_size + 2.5 * square_size
    targets = [(reward_x, 0, square_size / 2)]
    obstacles = [(i * square_size, 0, square_size / 2) for i in range(1, num_obstacles + 1)]
    return dict(shapes=_library(trapezoid, hexagon), obstacles=obstacles, targets=targets)


def bridge_setup(H=.8, num_stories=1, trapezoid=True, hexagon=False):
    """gym_env.py:46-61."""
    targets = [(0.5, 0, num_stories * H + H / 2)]
    obstacles = [(targets[0][0], 0., i * H + H / 2) for i in range(num_stories)]
    return dict(shapes=_library(trapezoid, hexagon), obstacles=obstacles, targets=targets)


def tower_setup(num_targets=3, targets=None):
    """gym_env.py:64-79."""
    if targets is None:
        targets = [(np.random.uniform(-4, 4), 0, np.random.uniform(0., 4)) for _ in range(num_targets)]
    return dict(shapes=[Shape(urdf_file='shapes/trapezoid.urdf', name="trapezoid")], obstacles=[], targets=targets)


def hard_tower_setup():
    """gym_env.py:82-88."""
    trapezoid = Shape(urdf_file='shapes/trapezoid.urdf', name="trapezoid")
    cube = Shape(urdf_file='shapes/cube1.urdf', name="cube", receiving_faces_2d=[0], target_faces_2d=[2])
    return dict(shapes=[trapezoid, cube], targets=[[0, 0, 0.5], [0, 0, 5.5]], obstacles=[[0, 0, 2.0]])


def connecting_setup():
    """gym_env.py:91-99."""
    rectangle = Shape(urdf_file='shapes/block.urdf', name="rectangle", receiving_faces_2d=[3], target_faces_2d=[0])
    cube = Shape(urdf_file='shapes/cube1.urdf', name="cube", receiving_faces_2d=[3], target_faces_2d=[1])
    targets = [[np.random.uniform(0.4, 0.6), 0, 0.175] for _ in range(3)]
    obstacles = [[np.random.uniform(0.4, 0.47), 0, np.random.uniform(0.025, 0.125)],
                 [np.random.uniform(0.53, 0.6), 0, np.random.uniform(0.025, 0.125)]]
    return dict(shapes=[rectangle, cube], obstacles=obstacles, targets=targets)


def tower_height_setup(tower_height=2, square_size=0.6):
    """The `--tower_height=k` task named by BASELINE.json (absent from the reference snapshot;
    definition: SURVEY.md section 8(d).3)."""
    obstacles = [(square_size, 0, i * square_size + square_size / 2) for i in range(tower_height - 1)]
    targets = [(square_size, 0, (tower_height - 1) * square_size + square_size / 2)]
    return dict(shapes=[Shape(urdf_file='shapes/trapezoid.urdf', name="trapezoid")], obstacles=obstacles,
                targets=targets)


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
    metadata = {"render_modes": ["human", "rgb_array"], "render_fps": 4}

    def __init__(self, reward_fct, shapes=None, obstacles=None, targets=None, render_mode=None, assembly_env=None,
                 restrict_2d=False, max_steps=None):
        self.blocks = []
        self.shapes = []
        self.obstacles = []
        self.targets = []
        self.reward_fct = reward_fct
        self.render_mode = render_mode
        self.restrict_2d = restrict_2d
        self.observation_space = None
        self.action_space = None
        self.action_history = None
        self.block_graph = None
        self.max_steps = max_steps
        if not restrict_2d:
            raise NotImplementedError
        if assembly_env is None:
            assembly_env = AssemblyEnv(render=render_mode == 'human')
        self.assembly_env = assembly_env
        assembly_env._gym = self
        self._core = None
        self._core_shapes = None
        self._marker = Shape(urdf_file='shapes/cube06.urdf')
        self.reset(shapes, obstacles, targets)

    # ------------------------------------------------------------------ GPU plumbing
    def _ensure_core(self):
        key = tuple((s.tables.urdf_file, tuple(s._target_faces_2d or ()), tuple(s._receiving_faces_2d or ()))
                    for s in self.shapes)
        if self._core is None or self._core_shapes != key:
            if self._core is not None:
                self._core.close()
            ae = self.assembly_env
            self._core = BatchedAssemblyGym(1, self.shapes, max_steps=self.max_steps, device=ae.device, mu=ae.mu,
                                            density=ae.density, bounds=ae.bounds, collision=ae.collision,
                                            collision_tol=ae.collision_tol)
            self._core_shapes = key
        return self._core

    def _sync_blocks(self):
        """Rebuild the host-side Block views from the device state."""
        blocks, n = self._core.get_state()
        views = []
        for i in range(int(n[0])):
            b = blocks[0][i]
            blk = Block(self.shapes[int(b["shape"])], [b["x"], 0.0, b["z"]], pose=(b["x"], b["z"], b["c"], b["s"]))
            blk.is_static = bool(b["is_static"])
            blk.object_id = i
            views.append(blk)
        self.assembly_env.blocks = views
        return views

    def _absorb(self, out):
        """bw_step_out -> state_info of the AssemblyEnv (assembly_env.py:307-324)."""
        ae = self.assembly_env
        stable = bool(out["stable"])
        if out["solver_status"] & 1:
            stable = None                                      # solver error -> None (stability.py:66-68)
        ae._state_info = {
            "last_block": ae.blocks[-1] if ae.blocks else None,
            "collision": bool(out["collision"]),
            # the step reports flags, not object ids: a non-empty list stands for "at least one"
            "collision_info": {"obstacles": [True] if out["collision_obstacle"] else [],
                               "blocks": [True] if out["collision_block"] else [],
                               "floor": bool(out["collision_floor"]), "bounding_box": bool(out["collision_boundary"])},
            "frozen_block": ae.frozen_block_index,
            "stable": stable if ae.stability else None,
            "stability_info": None if not (out["solver_status"] & 1) else dict(error="not converged"),
            "residual": float(out["residual"]),
        }
        self._last_out = out.copy()

    def _evaluate(self):
        self._core.evaluate()
        self._absorb(self._core.read_out()[0])

    def _reset_world(self):
        if self._core is not None:
            self._core.reset(dict(obstacles=self.obstacles, targets=self.targets))
            self.assembly_env.blocks = []
            self.assembly_env.obstacles = []
            self._evaluate()

    def _add_block(self, block):
        """AssemblyEnv.add_block (assembly_env.py:327-333) for an externally posed Block: the
        world is re-created with the extra block, supports kept."""
        core = self._ensure_core()
        blocks = self.assembly_env.blocks + [block]
        static = sum(1 << i for i, b in enumerate(blocks) if b.is_static)
        core.reset(dict(obstacles=self.obstacles, targets=self.targets,
                        blocks=[(b.pose[0], b.pose[1], b.pose[2], b.pose[3], self.shapes.index(b.shape)) for b in blocks]))
        core.set_static_mask(static)
        self._sync_blocks()
        for b, old in zip(self.assembly_env.blocks, blocks):
            b.is_static = old.is_static
        self._evaluate()
        return self.assembly_env.state_info

    def _set_static(self, index, value):
        blocks = self.assembly_env.blocks
        blocks[index].is_static = value
        self._core.set_static_mask(sum(1 << i for i, b in enumerate(blocks) if b.is_static))

    # ------------------------------------------------------------------ reference API
    def terminated(self, assembly_env):
        terminated = (not assembly_env.state_info['stable'] or assembly_env.state_info['collision']
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
        return [float(d) for d in self._last_out["distance_to_targets"][:self.num_targets]]

    def all_targets_reached(self):
        return len(self.targets_remaining) == 0

    def _pull_targets(self):
        core = self._core
        rem = np.zeros((1, L.BW_MAX_TARGETS), dtype=np.int8)
        rea = np.zeros((1, L.BW_MAX_TARGETS), dtype=np.int8)
        cnt = np.zeros((1, 2), dtype=np.int32)
        core._check(core.lib.bw_get_target_state(core.handle, rem.ctypes.data, rea.ctypes.data, cnt.ctypes.data))
        self.targets_remaining = [self.targets[i] for i in rem[0][:cnt[0][0]]]
        self.targets_reached = [self.targets[i] for i in rea[0][:cnt[0][1]]]

    def _get_obs(self):
        info = self.assembly_env.state_info
        return {
            'blocks': self.blocks,
            'stable': bool(info['stable']),
            'collision': bool(info['collision']),
            'collision_block': bool(info['collision_info']['blocks']),
            'collision_obstacle': bool(info['collision_info']['obstacles']),
            'collision_floor': bool(info['collision_info']['floor']),
            'collision_boundary': bool(info['collision_info']['bounding_box']),
            'frozen_block': self.assembly_env.frozen_block_index,
            'obstacles': self.obstacles,
            'obstacle_blocks': self.assembly_env.obstacles,
            'targets': self.targets,
            'targets_remaining': self.targets_remaining,
            'targets_reached': self.targets_reached,
            'distance_to_targets': self.distance_to_targets(),
        }

    def _get_info(self):
        return {'blocks_initial_state': None, 'blocks_final_state': None}

    def _query(self, action, xlim=None, ylim=None):
        core = self._core
        act = core.actions_array([action])
        blk = np.zeros(1, dtype=core.dt["block"])
        flags = np.zeros(1, dtype=np.uint8)
        xl = np.asarray(xlim, dtype=np.float64) if xlim is not None else None
        yl = np.asarray(ylim, dtype=np.float64) if ylim is not None else None
        core._check(core.lib.bw_query_placement_host(core.handle, act.ctypes.data,
                                                     xl.ctypes.data if xl is not None else None,
                                                     yl.ctypes.data if yl is not None else None,
                                                     blk.ctypes.data, flags.ctypes.data))
        if flags[0] & 1:
            raise IndexError(f"invalid action indices: {action}")
        return blk[0], int(flags[0])

    def create_block(self, action: Action):
        """gym_env.py:204-216 (placement computed by the CUDA library)."""
        b, _ = self._query(action)
        return Block(self.shapes[action.shape], [b["x"], 0.0, b["z"]], pose=(b["x"], b["z"], b["c"], b["s"]))

    def collision_on_action(self, action, xlim, ylim):
        """gym_env.py:304-323."""
        _, flags = self._query(action, xlim, ylim)
        return bool(flags & 4)

    def step(self, action: Action):
        """gym_env.py:218-253."""
        core = self._core
        core.step([action])
        out = core.read_out()[0]
        if out["error"] == 1:
            raise IndexError(f"invalid action indices: {action}")
        if out["error"] == 2:
            raise L.BridgesError("environment capacity exceeded (BW_MAX_BLOCKS / max_steps / BW_MAX_INTERFACES)")
        self.action_history.append(action)
        self.blocks = self._sync_blocks()
        new_index = len(self.blocks) - 1
        key = (action.target_block, action.target_face)
        self.block_graph.setdefault(key, []).append((new_index, action.face))
        self.block_graph[(new_index, action.face)] = [key]
        action.frozen = True                  # gym_env.py:238
        self._pull_targets()
        self._absorb(out)
        terminated, truncated = self.terminated(self.assembly_env)
        info = self._get_info()
        observation = self._get_obs()
        reward = self.reward_fct(self, observation, info)
        return observation, reward, terminated, truncated, info

    def reset(self, shapes=None, obstacles=None, targets=None, blocks=None):
        """gym_env.py:255-289."""
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
        self.targets_remaining = list(self.targets).copy()
        ae = self.assembly_env
        ae.blocks = []
        ae.obstacles = [Block(shape=self._marker, position=p) for p in self.obstacles]
        self._last_out = None
        if not self.shapes:
            return self._get_obs(), self._get_info()          # nothing to place yet (constructor without shapes)
        core = self._ensure_core()
        pre = []
        if blocks is not None:
            for b in blocks:                                  # 8-tuples x, y, z, qw, qx, qy, qz, shape
                blk = Block(self.shapes[b[-1]], b[:3], tuple(b[3:7]))
                pre.append((blk.pose[0], blk.pose[1], blk.pose[2], blk.pose[3], int(b[-1])))
        core.reset(dict(obstacles=self.obstacles, targets=self.targets, blocks=pre))
        self.blocks = self._sync_blocks() if pre else []
        self._evaluate()
        return self._get_obs(), self._get_info()

    @property
    def num_step(self):
        return len(self.action_history)

    def render(self):
        raise NotImplementedError("PyBullet rendering is not part of bridges_b200")

    def close(self):
        if self._core is not None:
            self._core.close()
            self._core = None

    def stabilities_freezing(self):
        """gym_env.py:325-333: both verdicts were produced by the step kernel."""
        return bool(self._last_out["stable"]), bool(self._last_out["stable_unfrozen"])
