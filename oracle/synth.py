"""Seeded random assemblies for parity tests and the CPU baseline.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Input generator of the
"synthetic stability sweep" (input 1): random valid
rollouts with shapes {trapezoid, hexagon, cube1}, 1..max_blocks blocks.
"""
import numpy as np

from .assembly_env import AssemblyEnv, Shape
from .gym_env import Action, AssemblyGym, sparse_reward

MUS = (0.3, 0.8, 2.0)


def sat_overlap(pa, pb, eps=1e-7):
    """Convex polygons overlap with penetration depth > eps (separating-axis test)."""
    for poly in (pa, pb):
        n = len(poly)
        for i in range(n):
            x0, z0 = poly[i]
            x1, z1 = poly[(i + 1) % n]
            ax, az = z1 - z0, -(x1 - x0)
            ln = (ax * ax + az * az) ** 0.5
            ax, az = ax / ln, az / ln
            a = [x * ax + z * az for x, z in pa]
            b = [x * ax + z * az for x, z in pb]
            if min(max(a) - min(b), max(b) - min(a)) <= eps:
                return False
    return True


def library():
    return [Shape(urdf_file="shapes/trapezoid.urdf", name="trapezoid"),
            Shape(urdf_file="shapes/hexagon.urdf", name="hexagon"),
            Shape(urdf_file="shapes/cube1.urdf", name="cube")]


def random_assembly(rng, shapes, max_blocks=15, xlim=(-3.0, 7.0), ylim=(0.0, 10.0), tries=40, scale=1.0, min_blocks=1):
    """Returns the list of Actions of one random valid rollout (no stability filter), min_blocks..max_blocks long
    (shorter when no further block can be placed).
    `scale` shrinks the offsets for the small shapes of the library (block.urdf, small_cube.urdf, ...)."""
    env = AssemblyGym(shapes=shapes, targets=[], obstacles=[], reward_fct=sparse_reward, restrict_2d=True,
                      assembly_env=AssemblyEnv(stability=None))
    n_blocks = int(rng.integers(min_blocks, max_blocks + 1))
    actions = []
    occupied = set()
    for k in range(n_blocks):
        for _ in range(tries):
            shape = int(rng.integers(len(shapes)))
            face = int(rng.integers(shapes[shape].num_faces_2d))
            if k == 0 or rng.random() < 0.2:
                action = Action(-1, 0, shape, face, float(rng.uniform(-2.0, 4.0)) * scale, 0.0)
            else:
                tb = int(rng.integers(k))
                tf = int(rng.integers(env.assembly_env.blocks[tb].num_faces_2d))
                if (tb, tf) in occupied:
                    continue
                action = Action(tb, tf, shape, face, float(rng.choice([0.0, 0.25, -0.25])) * scale, 0.0)
            if env.collision_on_action(action, xlim, ylim):
                continue
            block = env.create_block(action)
            if any(sat_overlap(block.polygon_2d, other.polygon_2d, eps=1e-7 * scale) for other in env.assembly_env.blocks):
                continue
            env.assembly_env.blocks.append(block)
            if action.target_block >= 0:
                occupied.add((action.target_block, action.target_face))
            occupied.add((k, face))
            actions.append(action)
            break
        else:
            break
    return actions


def replay(actions, shapes, mu, frozen_last, density=1.0):
    """Build the oracle AssemblyEnv holding the assembly (stability not evaluated)."""
    env = AssemblyGym(shapes=shapes, targets=[], obstacles=[], reward_fct=sparse_reward, restrict_2d=True,
                      assembly_env=AssemblyEnv(stability=None, mu=mu, density=density))
    for a in actions:
        env.assembly_env.blocks.append(env.create_block(a))
    for b in env.assembly_env.blocks:
        b.is_static = False
    if frozen_last and env.assembly_env.blocks:
        env.assembly_env.blocks[-1].is_static = True
    env.assembly_env._reset_cra_assembly()
    return env
