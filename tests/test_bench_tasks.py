"""bench.py's synthetic tasks are the reference's set-ups (gym_env.py:24-42) -- host logic, no GPU."""
import argparse

import bench


def _args(**kw):
    d = dict(task="tower", tower_height=2, num_obstacles=5, shapes="trapezoid,hexagon", envs=1024, max_steps=10)
    d.update(kw)
    return argparse.Namespace(**d)


def test_bridge_task_matches_the_adapter_and_the_oracle():
    from bridges_b200.envs.gym_env import horizontal_bridge_setup
    from oracle.gym_env import horizontal_bridge_setup as oracle_setup
    for n in (1, 5, 7):
        mine = bench.bridge_def(n)
        ref = horizontal_bridge_setup(num_obstacles=n, trapezoid=True, hexagon=True)
        orc = oracle_setup(num_obstacles=n, trapezoid=True, hexagon=True)
        assert mine["obstacles"] == ref["obstacles"] == orc["obstacles"]
        assert mine["targets"] == ref["targets"] == orc["targets"]


def test_task_spec_and_config_name_the_workload():
    tower = bench.task_spec(_args())
    assert tower["shapes"] == ["trapezoid"] and tower["obstacles"] == [(0.6, 0, 0.3)]
    assert tower["targets"] == [(0.6, 0, 0.6 + 0.3)]
    assert "configs[1]" in bench.config_dict(_args(), 1)["workload"]
    bridge = bench.task_spec(_args(task="bridge", max_steps=15))
    assert bridge["shapes"] == ["trapezoid", "hexagon"] and len(bridge["obstacles"]) == 5
    cfg = bench.config_dict(_args(task="bridge", max_steps=15), 8)
    assert "horizontal_bridge_setup(num_obstacles=5)" in cfg["workload"] and "x8" in cfg["parallelism"]


def test_cpu_arm_steps_both_tasks():
    """The CPU arm (restated reference env, oracle/) advances on the tower and on the bridge task."""
    for spec, max_steps in ((2, 10), (bench.task_spec(_args(task="bridge", max_steps=15)), 15)):
        steps, timed = bench._cpu_worker((1000, 20.0, spec, max_steps, 6, 1, 0.0))
        assert steps >= 6 and timed > 0.0
