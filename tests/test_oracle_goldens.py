"""The oracle against every golden vector the reference holds for the hot path (CPU only):
stored notebook outputs (tests/golden/notebook_goldens.json, harvested by make_goldens.py)
and the expected labels of utils/structures.py."""
import json
import os

import numpy as np
import pytest

from oracle import compas_lite as cl
from oracle import stability as st
from oracle.assembly_env import AssemblyEnv, Block, Shape
from oracle.gym_env import (Action, AssemblyGym, hard_tower_setup, horizontal_bridge_setup, sparse_reward)
from tests import fixtures_structures as FS
from tests import helpers as H

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "notebook_goldens.json")))


def test_hard_tower_notebook_run():
    g = GOLD["hard_tower"]
    env = AssemblyGym(**hard_tower_setup(), reward_fct=sparse_reward, restrict_2d=True, assembly_env=AssemblyEnv())
    for a, want in zip(g["actions"], g["steps"]):
        obs, reward, terminated, truncated, _ = env.step(Action(*a))
        assert obs["stable"] == want["stable"] and obs["collision"] == want["collision"]
        assert reward == want["reward"] and terminated == want["terminated"] and truncated is None
        assert len(obs["blocks"]) == want["n_blocks"]
        assert [list(t) for t in obs["targets_remaining"]] == want["targets_remaining"]
        assert [list(t) for t in obs["targets_reached"]] == want["targets_reached"]
        # bit-identical distances (16 digits in the notebook)
        assert obs["distance_to_targets"] == want["distance_to_targets"]


def test_horizontal_bridge_notebook_run():
    g = GOLD["horizontal_bridge_7_mu2"]
    env = AssemblyGym(**horizontal_bridge_setup(num_obstacles=g["num_obstacles"]), reward_fct=sparse_reward,
                      restrict_2d=True, assembly_env=AssemblyEnv(mu=g["mu"]))
    for a, want in zip(g["actions"], g["steps"]):
        obs, reward, terminated, truncated, _ = env.step(Action(*a))
        assert (obs["stable"], obs["collision"], len(obs["targets_reached"]), reward, terminated) == \
               (want["stable"], want["collision"], want["n_reached"], want["reward"], want["terminated"])


@pytest.mark.parametrize("offset", [0.8, 0.5])
def test_three_trapezoids(offset):
    g = GOLD["three_trapezoids"]
    env = AssemblyGym(shapes=[Shape(urdf_file="shapes/trapezoid.urdf")], targets=[], obstacles=[],
                      reward_fct=sparse_reward, restrict_2d=True, assembly_env=AssemblyEnv())
    ae = env.assembly_env
    flags = []
    for tb, tf, face, ox in g["placements"]:
        if tb == 0:
            ox = offset
        info = ae.add_block(env.create_block(Action(tb, tf, 0, face, ox, 0)))
        flags.append(info["stable"])
    assert flags[0] is True
    assert flags[1:] == g["stable_after_block2_and_3"]
    assert len(ae.cra_assembly.interfaces) == g["n_interfaces"]


def test_box_on_support_compressions():
    # CRA_Assembly.ipynb cells 2-4: Box(1,3,1) of density 1 resting on a fixed Box(4,2,1): the four
    # interface vertices carry 0.75 each = two 2-D contact points with 1.5 each
    support = Block(Shape(mesh=cl.mesh_from_box(4, 2, 1)), [0.0, 0.0, 0.0])
    free = Block(Shape(mesh=cl.mesh_from_box(1, 3, 1)), [0.0, 0.0, 1.0])
    support.is_static = True
    asm = st.CRAAssembly([[-5, -5, -1], [5, 5, 9]], [support, free])
    asm.interfaces = [i for i in asm.interfaces if i.a >= 0]      # the notebook has no floor
    assert len(asm.interfaces) == 1
    A, b = st.equilibrium_system(asm, 0.84, 1.0)
    f, y, r, status = st.min_norm_forces(A, b, 0.84)
    assert status == "feasible" and r < 1e-9
    assert np.allclose(f[0::2] / 2.0, GOLD["box_on_support"]["compressions"][:2], atol=1e-9)
    assert np.allclose(f[1::2], 0.0, atol=1e-9)


def _verdict(env, frozen_last):
    ae = env.assembly_env
    for blk in ae.blocks:
        blk.is_static = False
    if frozen_last:
        ae.blocks[-1].is_static = True
    ae._reset_cra_assembly()
    return st.is_stable_rbe(ae)[0]


@pytest.mark.parametrize("mu", [0.8, 0.3, 2.0])
def test_structures_labels(mu):
    checked = 0
    for name, mu_, fl, shapes, steps in FS.cases((mu,)):
        env = H.oracle_env(shapes, mu=mu_)
        env.assembly_env.stability_fct = lambda e: (None, None)
        for i, (a, expected) in enumerate(steps):
            env.step(Action(*a[:6]))
            got = _verdict(env, frozen_last=a[6])
            if (name, fl, i) in FS.KNOWN_LABEL_MISSES:
                assert got is True          # physically stable arch; label is the constant freeze_last
                continue
            assert got == bool(expected), (name, mu_, fl, i)
            checked += 1
    assert checked >= 20


def test_edge_less_rule_and_quirks():
    env = H.oracle_env(["cube"], mu=0.8)
    assert env.assembly_env.state_info["stable"] is True           # empty scene
    obs, reward, terminated, truncated, _ = env.step(Action(-1, 0, 0, 0, 0, 0.5))
    assert obs["stable"] is True                                    # single frozen floating block
    assert env.stabilities_freezing() == (True, False)              # released: free floating block
    assert truncated is None and obs["frozen_block"] is None
    assert terminated is True                                       # no targets: all_targets_reached()


def test_update_targets_skips_after_removal():
    # gym_env.py:162-168 removes from the list it iterates: the second of two targets hit by
    # the same block is skipped
    env = H.oracle_env(["cube"], targets=[(0.0, 0, 0.2), (0.1, 0, 0.3), (0.2, 0, 0.4)])
    obs, *_ = env.step(Action(-1, 0, 0, 0, 0.0, 0.0))
    assert [t[0] for t in obs["targets_reached"]] == [0.0, 0.2]
    assert [t[0] for t in obs["targets_remaining"]] == [0.1]
