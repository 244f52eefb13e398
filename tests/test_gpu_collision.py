"""collision_mode = 1 (`AssemblyEnv(pybullet_env=True)` of the reference, assembly_env.py:346-391): the
flags of `_check_collision` for the last block -- bounds test on the block position, penetration deeper
than tol = 0.005 against blocks / floor / obstacles -- are bit-exact against the oracle's polygon
penetration test; reward and termination follow gym_env.py:11-22,141-144."""
import numpy as np
import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu
BAND = (1e-9, 1e-4)
NAMES = ["trapezoid", "hexagon", "cube1"]


def _oracle(obstacles, targets, mu=0.8):
    from oracle.assembly_env import AssemblyEnv as OEnv
    from oracle.assembly_env import Shape as OShape
    from oracle.gym_env import AssemblyGym as OGym
    from oracle.gym_env import sparse_reward
    shapes = [OShape(urdf_file=H.URDF[n], name=n) for n in NAMES]
    return OGym(shapes=shapes, obstacles=list(obstacles), targets=list(targets), reward_fct=sparse_reward,
                restrict_2d=True, assembly_env=OEnv(mu=mu, pybullet_env=True))


def _random_action(rng, oenv):
    """placements that slide, sink, float and leave the bounds on purpose"""
    n = len(oenv.assembly_env.blocks)
    shape = int(rng.integers(len(NAMES)))
    face = int(rng.integers(oenv.shapes[shape].num_faces_2d))
    oy = float(rng.choice([0.0, 0.0, 0.0, -0.003, -0.0051, -0.02, -0.3, 0.05]))
    if n == 0 or rng.random() < 0.35:
        return (-1, 0, shape, face, float(rng.uniform(-3.6, 7.6)), oy)
    tb = int(rng.integers(n))
    tf = int(rng.integers(oenv.assembly_env.blocks[tb].num_faces_2d))
    return (tb, tf, shape, face, float(rng.choice([0.0, 0.0, 0.3, -0.3, 0.8, -1.1])), oy)


def test_collision_flags_reward_termination_bit_exact():
    from bridges_b200.envs.batched import BatchedAssemblyGym
    from oracle.gym_env import Action as OAction
    rng = np.random.default_rng(31)
    N, STEPS = 192, 7
    tasks, oenvs = [], []
    for e in range(N):
        k = int(rng.integers(0, 4))
        obstacles = [(float(rng.uniform(-2.5, 6.5)), 0, float(rng.choice([0.3, 0.3, 0.9, 1.5]))) for _ in range(k)]
        targets = [(float(rng.uniform(-2, 6)), 0, float(rng.uniform(0.2, 3.0)))]
        tasks.append(dict(obstacles=obstacles, targets=targets))
        oenvs.append(_oracle(obstacles, targets))
    env = BatchedAssemblyGym(N, [H.URDF[n] for n in NAMES], collision=True)
    env.reset(tasks)
    binary = __import__("torch").zeros((N, 6), device="cuda")
    seen = dict(collision=0, block=0, obstacle=0, floor=0, boundary=0, clean=0, steps=0)
    for k in range(STEPS):
        acts = [_random_action(rng, oenvs[e]) for e in range(N)]
        env.step(acts, binary=binary)
        out = env.read_out()
        feats = binary.cpu().numpy()
        for e in range(N):
            obs, reward, terminated, truncated, _ = oenvs[e].step(OAction(*acts[e]))
            o = out[e]
            flags = (bool(o["collision"]), bool(o["collision_block"]), bool(o["collision_obstacle"]),
                     bool(o["collision_floor"]), bool(o["collision_boundary"]))
            want = (obs["collision"], obs["collision_block"], obs["collision_obstacle"], obs["collision_floor"],
                    obs["collision_boundary"])
            assert flags == want, (e, k, acts[e], flags, want)
            assert list(feats[e][1:]) == [float(v) for v in want]            # get_state_features order
            for name, f in zip(("collision", "block", "obstacle", "floor", "boundary"), want):
                seen[name] += f
            seen["clean"] += not want[0]
            seen["steps"] += 1
            r_frozen, _ = H.residuals(oenvs[e])
            if r_frozen is not None and BAND[0] < r_frozen < BAND[1]:
                continue
            assert bool(o["stable"]) == bool(obs["stable"]), (e, k)
            assert float(o["reward"]) == float(reward) and bool(o["terminated"]) == bool(terminated), (e, k)
            if want[0]:
                assert float(o["reward"]) == -1.0 and bool(o["terminated"])
    # every flag and the collision-free case are exercised
    assert min(seen[k] for k in ("block", "obstacle", "floor", "boundary")) >= 20 and seen["clean"] >= 200, seen


def test_collision_mode_off_is_constant_false():
    """assembly_env.py:310-312: without a physics client every flag is False whatever the placement."""
    from bridges_b200.envs.batched import BatchedAssemblyGym
    env = BatchedAssemblyGym(2, [H.URDF["cube1"]])
    env.reset(dict(obstacles=[(0.0, 0, 0.3)], targets=[(3.0, 0, 0.5)]))
    env.step([(-1, 0, 0, 0, 0.0, -0.4), (-1, 0, 0, 0, 9.0, 0.0)])      # inside the obstacle and the floor; out of bounds
    out = env.read_out()
    for name in ("collision", "collision_block", "collision_obstacle", "collision_floor", "collision_boundary"):
        assert not out[name].any()


def test_dropin_pybullet_env_flag_enables_collisions():
    from bridges_b200.envs.assembly_env import AssemblyEnv, Shape
    from bridges_b200.envs.gym_env import Action, AssemblyGym, sparse_reward
    gym = AssemblyGym(reward_fct=sparse_reward, shapes=[Shape(urdf_file="shapes/cube1.urdf")], obstacles=[(0.6, 0, 0.3)],
                      targets=[(3.0, 0, 0.5)], restrict_2d=True, assembly_env=AssemblyEnv(pybullet_env=True))
    obs, reward, terminated, truncated, _ = gym.step(Action(-1, 0, 0, 0, 0.0, 0.0))     # overlaps the obstacle cube
    assert obs["collision"] and obs["collision_obstacle"] and not obs["collision_floor"]
    assert reward == -1 and terminated
    assert gym.assembly_env.state_info["collision_info"]["obstacles"]
    gym.reset()
    obs, reward, terminated, *_ = gym.step(Action(-1, 0, 0, 0, 2.5, 0.0))               # free standing, reaches the target
    assert not obs["collision"] and reward == 1 and terminated
