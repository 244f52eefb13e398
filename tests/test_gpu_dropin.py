"""The single-environment drop-in classes (same names and call sequence as the reference's
AssemblyGym / AssemblyEnv / render_blocks_2d / generate_actions / filter_actions) against the
notebook goldens and the oracle."""
import json
import os

import numpy as np
import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "notebook_goldens.json")))


def test_hard_tower_notebook_through_dropin():
    from bridges_b200.envs.assembly_env import AssemblyEnv
    from bridges_b200.envs.gym_env import Action, AssemblyGym, hard_tower_setup, sparse_reward
    g = GOLD["hard_tower"]
    env = AssemblyGym(**hard_tower_setup(), reward_fct=sparse_reward, restrict_2d=True,
                      assembly_env=AssemblyEnv(render=False))
    for a, want in zip(g["actions"], g["steps"]):
        action = Action(*a)
        obs, reward, terminated, truncated, info = env.step(action)
        assert action.frozen is True                                 # caller's Action is mutated (quirk 1)
        assert obs["stable"] == want["stable"] and obs["collision"] == want["collision"]
        assert reward == want["reward"] and terminated == want["terminated"] and truncated is None
        assert len(obs["blocks"]) == want["n_blocks"]
        assert [list(t) for t in obs["targets_remaining"]] == want["targets_remaining"]
        assert [list(t) for t in obs["targets_reached"]] == want["targets_reached"]
        assert obs["distance_to_targets"] == want["distance_to_targets"]
        assert info == {"blocks_initial_state": None, "blocks_final_state": None}
    assert env.stabilities_freezing() == (True, True)
    env.close()


def test_rollout_api_against_oracle():
    import torch
    from bridges_b200.envs.assembly_env import AssemblyEnv
    from bridges_b200.envs.gym_env import Action, AssemblyGym, horizontal_bridge_setup, sparse_reward
    from bridges_b200.utils.actions import (filter_actions, generate_actions, get_action_features,
                                            get_state_features, get_task_features)
    from bridges_b200.utils.rendering import render_blocks_2d
    from oracle import actions as oact
    from oracle import features as ofeat
    from oracle.gym_env import Action as OAction
    from oracle.rendering import render_blocks_2d as o_render
    xg = np.linspace(-2, 0, 10)
    env = AssemblyGym(reward_fct=sparse_reward, max_steps=10, restrict_2d=True, assembly_env=AssemblyEnv(mu=2.0))
    obs, info = env.reset(**horizontal_bridge_setup(num_obstacles=7))
    oenv = H.oracle_env(["trapezoid"], [(i * 0.6, 0, 0.3) for i in range(1, 8)], [(7 * 0.6 + 1.5, 0, 0.3)],
                        mu=2.0, max_steps=10)
    oobs, _ = oenv.reset()
    reward_f, obstacle_f = get_task_features(obs, env=env)
    o_reward_f, o_obstacle_f = ofeat.get_task_features(oobs, H.XLIM, H.YLIM, H.IMG)
    assert np.array_equal(obstacle_f.cpu().numpy(), o_obstacle_f)
    assert np.allclose(reward_f.cpu().numpy(), o_reward_f, rtol=1e-5, atol=1e-7)
    for a in GOLD["horizontal_bridge_7_mu2"]["actions"][:5]:
        # candidate stage exactly as rollout_episode does it (successor_dqn.py:375-377)
        block_f, binary_f = get_state_features(obs, env=env)
        actions = [*generate_actions(env, x_discr_ground=xg, offset_values=[0.])]
        feats = get_action_features(env, actions)
        kept, kept_feats = filter_actions(env, actions, feats, block_features=block_f, obstacle_features=obstacle_f)
        ocands = [*oact.generate_actions(oenv, xg, [0.])]
        ocf = ofeat.get_action_features(oenv, ocands, H.XLIM, H.YLIM, H.IMG)
        obf, _ = ofeat.get_state_features(oobs, H.XLIM, H.YLIM, H.IMG)
        okept, okept_f, _ = oact.filter_actions(oenv, ocands, ocf, obf, o_obstacle_f, H.XLIM, H.YLIM)
        assert [(x.target_block, x.target_face, x.shape, x.face, x.offset_x) for x in kept] == \
               [(x.target_block, x.target_face, x.shape, x.face, x.offset_x) for x in okept]
        assert np.array_equal(kept_feats.cpu().numpy(), okept_f)
        # create_block / collision_on_action / render_blocks_2d one by one, as the reference's loops do
        for cand, ocand in list(zip(actions, ocands))[::7]:
            blk, oblk = env.create_block(cand), oenv.create_block(ocand)
            assert blk.pose == oblk.pose
            assert [tuple(v) for v in blk.vertices_2d] == [tuple(p) for p in oblk.polygon_2d]
            assert env.collision_on_action(cand, H.XLIM, H.YLIM) == oenv.collision_on_action(ocand, H.XLIM, H.YLIM)
            assert np.array_equal(render_blocks_2d([blk], H.XLIM, H.YLIM, H.IMG), o_render([oblk], H.XLIM, H.YLIM, H.IMG))
        obs, reward, terminated, truncated, _ = env.step(Action(*a))
        oobs, oreward, oterm, otrunc, _ = oenv.step(OAction(*a))
        assert (reward, terminated, bool(truncated)) == (oreward, oterm, bool(otrunc))
        assert env.stabilities_freezing() == oenv.stabilities_freezing()
        assert np.array_equal(render_blocks_2d(obs["blocks"], H.XLIM, H.YLIM, H.IMG),
                              o_render(oobs["blocks"], H.XLIM, H.YLIM, H.IMG))
        assert sorted(env.block_graph.items()) == sorted(oenv.block_graph.items())
    env.close()


def test_reset_with_preplaced_blocks_and_freeze_api():
    from bridges_b200.envs.assembly_env import AssemblyEnv, Shape
    from bridges_b200.envs.gym_env import AssemblyGym, sparse_reward
    from oracle.assembly_env import AssemblyEnv as OEnv
    from oracle.assembly_env import Shape as OShape
    from oracle.gym_env import AssemblyGym as OGym
    from oracle.gym_env import sparse_reward as o_reward
    blocks = [(0.0, 0.0, 0.5, 1.0, 0.0, 0.0, 0.0, 0), (0.7, 0.0, 1.5, 1.0, 0.0, 0.0, 0.0, 0)]   # second cube overhangs
    env = AssemblyGym(reward_fct=sparse_reward, restrict_2d=True, assembly_env=AssemblyEnv())
    obs, _ = env.reset(shapes=[Shape(urdf_file="shapes/cube1.urdf")], obstacles=[], targets=[(0.7, 0, 1.5)],
                       blocks=blocks)
    oenv = OGym(reward_fct=o_reward, restrict_2d=True, assembly_env=OEnv())
    oobs, _ = oenv.reset(shapes=[OShape(urdf_file="shapes/cube1.urdf")], obstacles=[], targets=[(0.7, 0, 1.5)],
                         blocks=blocks)
    assert obs["stable"] == oobs["stable"] is False
    assert obs["distance_to_targets"] == oobs["distance_to_targets"]
    # freezing the overhanging cube makes the assembly stable (AssemblyEnv.freeze_block + _update_state_info)
    env.assembly_env.freeze_block(1)
    env.assembly_env._update_state_info()
    oenv.assembly_env.freeze_block(1)
    oenv.assembly_env._update_state_info()
    assert env.assembly_env.is_stable() == oenv.assembly_env.is_stable() is True
    env.assembly_env.unfreeze_block(1)
    env.assembly_env._update_state_info()
    assert env.assembly_env.is_stable() is False
    env.close()


def test_errors_mirror_reference():
    from bridges_b200.envs.assembly_env import AssemblyEnv, Shape
    from bridges_b200.envs.gym_env import Action, AssemblyGym, sparse_reward
    with pytest.raises(NotImplementedError):
        AssemblyGym(reward_fct=sparse_reward, restrict_2d=False)          # gym_env.py:131-133
    with pytest.raises(FileNotFoundError):
        Shape(urdf_file="shapes/nope.urdf")                               # assembly_env.py:58-59
    env = AssemblyGym(reward_fct=sparse_reward, restrict_2d=True, shapes=[Shape(urdf_file="shapes/cube.urdf")],
                      obstacles=[], targets=[], assembly_env=AssemblyEnv())
    with pytest.raises(IndexError):
        env.step(Action(3, 0, 0, 0))                                      # blocks[3] does not exist
    env.close()
