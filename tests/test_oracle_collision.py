"""Oracle restatement of `_check_collision` (assembly_env.py:346-391): known-answer cases of the polygon
penetration test that replaces Bullet's contact distances (tol = 0.005)."""
import math

from oracle.assembly_env import AssemblyEnv, Block, Shape, polygon_separation
from oracle.gym_env import Action, AssemblyGym, sparse_reward


def _cube(x, z, c=1.0, s=0.0):
    return Block(Shape(urdf_file="shapes/cube1.urdf"), [x, 0.0, z], pose=(x, z, c, s))


def test_polygon_separation_known_answers():
    a = _cube(0.0, 0.5)
    assert polygon_separation(a, _cube(1.0, 0.5)) == 0.0                  # touching faces
    assert polygon_separation(a, _cube(1.5, 0.5)) == 0.5                  # gap
    assert polygon_separation(a, _cube(0.75, 0.5)) == -0.25               # overlap depth
    assert polygon_separation(a, _cube(0.75, 0.9)) == -0.25               # min translation axis is x
    # a cube turned by 45 degrees with its corner 0.1 inside the right face of `a`
    r = math.sqrt(0.5)
    b = _cube(0.5 + r - 0.1, 0.5, r, r)
    assert abs(polygon_separation(a, b) + 0.1) < 1e-12
    assert polygon_separation(a, b) == polygon_separation(b, a)


def test_check_collision_flags_and_tolerance():
    def run(ox, oy, obstacles=((0.6, 0, 0.3),)):
        gym = AssemblyGym(sparse_reward, shapes=[Shape(urdf_file="shapes/cube1.urdf")], obstacles=list(obstacles),
                          targets=[(9.0, 0, 0.5)], restrict_2d=True, assembly_env=AssemblyEnv(pybullet_env=True))
        obs, reward, terminated, _, _ = gym.step(Action(-1, 0, 0, 0, ox, oy))
        return obs, reward, terminated
    obs, reward, terminated = run(0.0, 0.0)
    assert obs["collision"] and obs["collision_obstacle"] and not obs["collision_block"] and reward == -1 and terminated
    obs, reward, terminated = run(1.4, 0.0)                                # exactly beside the obstacle: touching
    assert not obs["collision"] and not terminated
    assert not run(3.0, -0.0049)[0]["collision_floor"]                     # only penetration deeper than tol counts
    assert run(3.0, -0.0051)[0]["collision_floor"]
    assert run(7.2, 0.0, obstacles=())[0]["collision_boundary"] and not run(6.9, 0.0, obstacles=())[0]["collision_boundary"]
    # default configuration: no physics client, constant False (assembly_env.py:310-312)
    gym = AssemblyGym(sparse_reward, shapes=[Shape(urdf_file="shapes/cube1.urdf")], obstacles=[(0.6, 0, 0.3)],
                      targets=[(9.0, 0, 0.5)], restrict_2d=True)
    assert not gym.step(Action(-1, 0, 0, 0, 0.0, -0.3))[0]["collision"]
