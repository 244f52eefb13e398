"""The reference's own trainer on the drop-in.

tests/golden/reference_rollout_trace.json.gz is the log of every call the reference's UNMODIFIED `rollout_episode`
(robotoddler/training/successor_dqn.py:365-475, with `EpsilonGreedy`, `generate_actions` / `filter_actions` of
robotoddler/utils/actions.py and `SuccessorMLP` of models/cv.py:76-105; `train_policy_net` :157-277 then consumed
the transitions) made into the environment API during three episodes, recorded against the CPU oracle by
tests/golden/make_reference_rollout.py.  The GPU test issues the same calls to the drop-in classes of bridges_b200
(`AssemblyGym`, `Action`, `Shape`, `Block`, `render_blocks_2d`) and demands the same answers: observations, rewards,
termination, block_graph, posed candidate blocks bit for bit, collision_on_action flags, and every rendered
64 x 64 image.  Everything the trainer derives from them is then the same as well.

Where /root/reference exists (the build container; never on the GPU box) two CPU tests also check that the
reference's training script imports cleanly with `assembly_gym` bound to bridges_b200, and that the committed trace
is what the generator produces today."""
import gzip
import hashlib
import json
import os
import sys
import types

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TRACE = os.path.join(ROOT, "tests", "golden", "reference_rollout_trace.json.gz")
REFERENCE = "/root/reference"


def _load():
    with gzip.open(TRACE, "rb") as fh:
        return json.loads(fh.read().decode())


def _f(h):
    return float.fromhex(h)


def test_trace_fixture_is_well_formed():
    doc = _load()
    ops = [c["op"] for c in doc["calls"]]
    assert ops.count("reset") == 3 and ops.count("step") == sum(len(e) for e in doc["transitions"]) == 14
    assert ops.count("stabilities_freezing") == ops.count("step")
    assert ops.count("render_blocks_2d") > 1000 and ops.count("create_block") > 1000 and ops.count("collision_on_action") > 1000
    # the reference's learner consumed the transitions (two Adam steps per episode on the SuccessorMLP)
    assert len(doc["train_losses"]) == 3 and all(len(l) == 2 and all(np.isfinite(l)) for l in doc["train_losses"])
    # every episode ends with done, and only there
    for ep in doc["transitions"]:
        assert [t["done"] for t in ep] == [False] * (len(ep) - 1) + [True]


@pytest.mark.gpu
def test_reference_call_trace_replayed_on_the_drop_in():
    from bridges_b200.envs.assembly_env import AssemblyEnv, Block, Shape
    from bridges_b200.envs.gym_env import Action, AssemblyGym, sparse_reward
    from bridges_b200.utils.rendering import render_blocks_2d
    doc = _load()
    env = AssemblyGym(reward_fct=sparse_reward, max_steps=doc["config"]["max_steps"], restrict_2d=True,
                      assembly_env=AssemblyEnv(render=False))            # successor_dqn.py:695
    marker = Shape(urdf_file="shapes/cube06.urdf")                         # get_task_features, successor_dqn.py:73
    by_name = {}

    def action_of(a):
        return Action(a[0], a[1], a[2], a[3], _f(a[4]), offset_y=_f(a[5]))

    def same_obs(obs, want, tag):
        assert [[b.name, [v.hex() for v in b.pose]] for b in obs["blocks"]] == want["blocks"], tag      # bit-exact poses
        for k in ("stable", "collision", "collision_block", "collision_obstacle", "collision_floor", "collision_boundary"):
            assert bool(obs[k]) == want[k], (tag, k)
        assert [[float(c).hex() for c in t] for t in obs["targets_remaining"]] == want["targets_remaining"], tag
        assert [[float(c).hex() for c in t] for t in obs["targets_reached"]] == want["targets_reached"], tag
        assert [float(d).hex() for d in obs["distance_to_targets"]] == want["distance_to_targets"], tag
        assert len(obs["obstacle_blocks"]) == want["n_obstacle_blocks"], tag

    counts = {}
    for i, c in enumerate(doc["calls"]):
        op, tag = c["op"], (i, c["op"])
        counts[op] = counts.get(op, 0) + 1
        if op == "reset":
            shapes = [Shape(urdf_file=u, name=n) for u, n in c["shapes"]]
            by_name = {s.name: s for s in shapes}
            obs, info = env.reset(shapes=shapes, obstacles=[[_f(v) for v in p] for p in c["obstacles"]],
                                  targets=[[_f(v) for v in p] for p in c["targets"]])
            same_obs(obs, c["obs"], tag)
        elif op == "step":
            a = action_of(c["action"])
            obs, reward, terminated, truncated, info = env.step(a)
            assert a.frozen is True                                        # gym_env.py:238 mutates the caller's Action
            same_obs(obs, c["obs"], tag)
            assert float(reward) == c["reward"] and bool(terminated) == c["terminated"], tag
            assert (bool(truncated) if truncated is not None else None) == c["truncated"], tag
            graph = sorted([list(k), [list(v) for v in vs]] for k, vs in env.block_graph.items())
            assert graph == c["block_graph"], tag
        elif op == "stabilities_freezing":
            assert [bool(v) for v in env.stabilities_freezing()] == c["result"], tag
        elif op == "create_block":
            b = env.create_block(action_of(c["action"]))
            assert [b.name, [v.hex() for v in b.pose]] == c["block"], tag
        elif op == "collision_on_action":
            got = env.collision_on_action(action_of(c["action"]), tuple(_f(v) for v in c["xlim"]), tuple(_f(v) for v in c["ylim"]))
            assert bool(got) == c["result"], tag
        elif op == "render_blocks_2d":
            blocks = []
            for name, pose in c["blocks"]:
                shape = by_name.get(name, marker)                          # "" = the cube06 marker of targets / obstacles
                p = [_f(v) for v in pose]
                blocks.append(Block(shape, [p[0], 0.0, p[1]], pose=p))
            img = render_blocks_2d(blocks, xlim=tuple(_f(v) for v in c["xlim"]), ylim=tuple(_f(v) for v in c["ylim"]),
                                   img_size=tuple(c["img_size"]))
            assert img.shape == tuple(c["img_size"]) and int(img.sum()) == c["pixels"], tag
            assert hashlib.sha1(np.packbits(img, axis=None).tobytes()).hexdigest() == c["sha1"], tag
    assert counts["step"] == 14 and counts["render_blocks_2d"] > 1000


@pytest.mark.gpu
def test_contains_2d_and_other_image_sizes():
    """Shape.contains_2d (assembly_env.py:126-137) and render_blocks_2d at a size other than 64 x 64
    (rendering.py:105-113) against the oracle, bit for bit."""
    from bridges_b200.envs.assembly_env import Block, Shape
    from bridges_b200.utils.rendering import render_blocks_2d
    from oracle.assembly_env import Block as OBlock
    from oracle.assembly_env import Shape as OShape
    from oracle.rendering import render_blocks_2d as o_render
    rng = np.random.default_rng(5)
    pts = rng.uniform(-2.0, 3.0, size=(4000, 2))
    poses = [(0.3, 0.36, 1.0, 0.0), (1.1, 1.2, 0.5000000126183913, 0.8660253965223742), (-0.4, 2.0, -1.0, 1.2e-16)]
    for urdf in ("shapes/trapezoid.urdf", "shapes/hexagon.urdf"):
        s, o = Shape(urdf_file=urdf), OShape(urdf_file=urdf)
        assert np.array_equal(s.contains_2d(pts), o.contains_2d(pts))
        for pose in poses:
            b, ob = Block(s, [pose[0], 0, pose[1]], pose=pose), OBlock(o, [pose[0], 0, pose[1]], pose=pose)
            on_edges = np.array([p for p in ob.polygon_2d] + [c for c in ob.face_centers_2d])    # knife-edge points too
            q = np.concatenate([pts, on_edges])
            assert np.array_equal(b.contains_2d(q), ob.contains_2d(q)), (urdf, pose)
    s, o = Shape(urdf_file="shapes/trapezoid.urdf"), OShape(urdf_file="shapes/trapezoid.urdf")
    blocks = [Block(s, [p[0], 0, p[1]], pose=p) for p in poses]
    oblocks = [OBlock(o, [p[0], 0, p[1]], pose=p) for p in poses]
    for size in ((32, 32), (96, 96), (128, 128)):
        assert np.array_equal(render_blocks_2d(blocks, (-3, 7), (0., 10), size), o_render(oblocks, (-3, 7), (0., 10), size)), size


needs_reference = pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "robotoddler")),
                                     reason="/root/reference only exists in the build container")


@needs_reference
def test_reference_training_script_imports_against_the_drop_in():
    """`assembly_gym` bound to bridges_b200: every name successor_dqn.py and robotoddler/utils/actions.py import
    from the environment package resolves (no environment is constructed: that needs the GPU)."""
    import bridges_b200.envs.assembly_env as d_ae
    import bridges_b200.envs.gym_env as d_gym
    import bridges_b200.utils.rendering as d_rend
    saved = dict(sys.modules)
    saved_path = list(sys.path)
    try:
        for name, m in (("assembly_gym", types.ModuleType("assembly_gym")), ("assembly_gym.envs", types.ModuleType("assembly_gym.envs")),
                        ("assembly_gym.utils", types.ModuleType("assembly_gym.utils")), ("assembly_gym.envs.gym_env", d_gym),
                        ("assembly_gym.envs.assembly_env", d_ae), ("assembly_gym.utils.rendering", d_rend),
                        ("aim", types.ModuleType("aim")), ("wandb", types.ModuleType("wandb"))):
            sys.modules[name] = m
        plt = types.ModuleType("matplotlib.pyplot")
        mpl = types.ModuleType("matplotlib")
        mpl.pyplot = plt
        sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, plt
        for k in [k for k in sys.modules if k == "robotoddler" or k.startswith("robotoddler.")]:
            del sys.modules[k]
        sys.path.insert(0, REFERENCE)
        from robotoddler.training import successor_dqn as sdqn
        assert sdqn.AssemblyGym is d_gym.AssemblyGym and sdqn.AssemblyEnv is d_ae.AssemblyEnv
        assert sdqn.render_blocks_2d is d_rend.render_blocks_2d and sdqn.horizontal_bridge_setup is d_gym.horizontal_bridge_setup
        assert callable(sdqn.rollout_episode) and callable(sdqn.train_policy_net)
        from robotoddler.utils import actions as ract
        assert ract.Action is d_gym.Action
    finally:
        sys.path[:] = saved_path
        for k in [k for k in sys.modules if k not in saved]:
            del sys.modules[k]
        sys.modules.update(saved)


@needs_reference
def test_trace_fixture_is_what_the_generator_produces(tmp_path):
    import subprocess
    out = tmp_path / "trace.json.gz"
    subprocess.run([sys.executable, os.path.join(ROOT, "tests", "golden", "make_reference_rollout.py"), str(out)], check=True,
                   capture_output=True, timeout=600)
    with gzip.open(out, "rb") as fh:
        fresh = json.loads(fh.read().decode())
    want = _load()
    assert fresh["calls"] == want["calls"] and fresh["transitions"] == want["transitions"]
