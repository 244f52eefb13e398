"""bridges_b200: B200-native batched assembly_gym environment step.

Drop-in for the env side of syghmon/bridges-with-reinforcement-learning:

    from bridges_b200.envs.gym_env import AssemblyGym, Action, sparse_reward, horizontal_bridge_setup
    from bridges_b200.envs.assembly_env import AssemblyEnv, Shape, Block
    from bridges_b200.utils.rendering import render_blocks_2d
    from bridges_b200.utils.actions import generate_actions, filter_actions

All numerics run in `libbridges_b200.so` (hand-written sm_100a CUDA behind the C ABI of
include/bridges_b200.h).  There is no CPU fallback: creating an environment without the
library or without a GPU raises.
"""
__version__ = "0.1.0"
