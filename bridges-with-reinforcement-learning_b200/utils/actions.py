"""Drop-in `generate_actions` / `filter_actions` (robotoddler/utils/actions.py:7-82) and the
feature functions of robotoddler/training/successor_dqn.py:47-94, served by the candidate
kernel (one launch enumerates, rasters and filters every candidate)."""
import numpy as np
import torch

from ..envs.gym_env import Action


def generate_actions(gym, x_discr_ground, offset_values=None, max_angle_rad=2 * np.pi + 0.1, max_blocks_per_face=1,
                     include_frozen=False, x_block_offset=None, amax=1024):
    """actions.py:7-52.  Yields Action objects in the reference's order."""
    if include_frozen:
        raise NotImplementedError
    if max_blocks_per_face != 1 or (max_angle_rad is not None and max_angle_rad < np.pi):
        raise NotImplementedError("only the configuration used by successor_dqn.py is implemented")
    core = gym._core
    c = core.enumerate_actions(x_discr_ground, offset_values if offset_values is not None else [0.], amax=amax)
    core.sync()
    n = int(c["n"][0].item())
    cand = c["cand"].cpu().numpy().view(core.dt["action"]).reshape(1, amax)[0][:n]
    gym._last_candidates = dict(n=n, valid=c["valid"][0, :n].clone(), bits=c["bits"][0, :n].clone())
    for a in cand:
        yield Action(int(a["target_block"]), int(a["target_face"]), int(a["shape"]), int(a["face"]),
                     float(a["offset_x"]), offset_y=float(a["offset_y"]))


def get_action_features(env, actions, xlim=(0, 1), ylim=(0, 1), img_size=(64, 64), device=None):
    """successor_dqn.py:88-94 for the candidates of the last `generate_actions(env, ...)` call."""
    last = getattr(env, "_last_candidates", None)
    if last is None or last["n"] != len(actions):
        raise RuntimeError("call generate_actions(env, ...) first; features are produced with the candidates")
    img = env._core.expand_bits(last["bits"].contiguous())
    return img if device is None else img.to(device)


def filter_actions(gym_env, available_actions, action_features, block_features=None, obstacle_features=None,
                   xlim=None, ylim=None):
    """actions.py:71-82: the mask was computed together with the candidates."""
    last = getattr(gym_env, "_last_candidates", None)
    if last is None or last["n"] != len(available_actions):
        raise RuntimeError("call generate_actions(gym_env, ...) first")
    mask = last["valid"].bool()
    keep = mask.cpu().numpy()
    reduced = [a for a, k in zip(available_actions, keep) if k]
    return reduced, action_features[mask.to(action_features.device)]


def get_state_features(observation, xlim=(0, 1), ylim=(0, 1), img_size=(64, 64), device=None, env=None):
    """successor_dqn.py:47-64.  With `env=` the tensors come straight from the device state."""
    if env is None:
        from .rendering import render_blocks_2d
        image = torch.from_numpy(render_blocks_2d(observation['blocks'], xlim, ylim, img_size).astype(np.float32))
        binary = torch.tensor([observation[k] for k in ('stable', 'collision', 'collision_block',
                                                        'collision_obstacle', 'collision_floor',
                                                        'collision_boundary')], dtype=torch.float32)
        return image.unsqueeze(0).to(device), binary.to(device)
    feats = env._core.observe(block=True, binary=True)
    return feats["block"][0].to(device), feats["binary"][0].to(device)


def get_task_features(obs, xlim=(0, 1), ylim=(0, 1), img_size=(64, 64), device=None, env=None):
    """successor_dqn.py:67-85 (needs `env=`: the images were rendered by reset)."""
    if env is None:
        raise RuntimeError("pass env=<AssemblyGym>: the task images live with the device state")
    feats = env._core.observe(block=False, binary=False, obstacle=True, reward=True)
    return feats["reward"][0].to(device), feats["obstacle"][0].to(device)
