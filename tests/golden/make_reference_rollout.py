"""Generates tests/golden/reference_rollout_trace.json.gz: the reference's UNMODIFIED `rollout_episode`
(robotoddler/training/successor_dqn.py:365-475), `EpsilonGreedy`, `generate_actions` / `filter_actions`
(robotoddler/utils/actions.py), `SuccessorMLP` (robotoddler/models/cv.py:76-105), `ReplayBuffer` and
`train_policy_net` (successor_dqn.py:157-277) are imported from /root/reference and run for a few episodes with the
`assembly_gym` package name bound to a RECORDING view of this repository's CPU oracle (oracle/, which restates
assembly_gym: the real package needs compas / compas_cra / pyomo / ipopt, none installable here).  Every call the
reference's code makes into the environment API -- reset, step, stabilities_freezing, create_block,
collision_on_action, render_blocks_2d -- is logged with its arguments and its result.

tests/test_reference_rollout.py (GPU) replays that call sequence against the drop-in classes of bridges_b200 and
demands identical results: whatever the reference's trainer computes from them (features, transitions, losses) is
then identical too.  /root/reference exists only in the build container; the trace is the part that travels.

    python tests/golden/make_reference_rollout.py            # rewrites the fixture (CPU only, ~1 minute)
"""
import hashlib
import json
import os
import random
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REFERENCE = "/root/reference"
sys.path.insert(0, ROOT)

from oracle import assembly_env as oae          # noqa: E402
from oracle import gym_env as ogym              # noqa: E402
from oracle import rendering as orend           # noqa: E402

TRACE = []


def fx(v):
    """exact, portable float"""
    return float(v).hex()


def act(a):
    return [int(a.target_block), int(a.target_face), int(a.shape), int(a.face), fx(a.offset_x), fx(a.offset_y)]


def blk(b):
    return [b.name, [fx(v) for v in b.pose]]


def bits_digest(img):
    return hashlib.sha1(np.packbits(np.asarray(img, dtype=bool), axis=None).tobytes()).hexdigest()


def obs_summary(obs):
    return dict(blocks=[blk(b) for b in obs["blocks"]], stable=bool(obs["stable"]), collision=bool(obs["collision"]),
                collision_block=bool(obs["collision_block"]), collision_obstacle=bool(obs["collision_obstacle"]),
                collision_floor=bool(obs["collision_floor"]), collision_boundary=bool(obs["collision_boundary"]),
                targets_remaining=[[fx(c) for c in t] for t in obs["targets_remaining"]],
                targets_reached=[[fx(c) for c in t] for t in obs["targets_reached"]],
                distance_to_targets=[fx(d) for d in obs["distance_to_targets"]],
                n_obstacle_blocks=len(obs["obstacle_blocks"]))


class RecordingGym(ogym.AssemblyGym):
    def reset(self, shapes=None, obstacles=None, targets=None, blocks=None):
        out = super().reset(shapes=shapes, obstacles=obstacles, targets=targets, blocks=blocks)
        if shapes is not None:              # (the constructor's own reset has nothing to place yet)
            TRACE.append(dict(op="reset", shapes=[[s.urdf_file, s.name] for s in shapes],
                              obstacles=[[fx(c) for c in p] for p in obstacles], targets=[[fx(c) for c in p] for p in targets],
                              obs=obs_summary(out[0])))
        return out

    def step(self, action):
        a = act(action)
        obs, reward, terminated, truncated, info = super().step(action)
        TRACE.append(dict(op="step", action=a, obs=obs_summary(obs), reward=float(reward), terminated=bool(terminated),
                          truncated=bool(truncated) if truncated is not None else None,
                          block_graph=sorted([list(k), [list(v) for v in vs]] for k, vs in self.block_graph.items())))
        return obs, reward, terminated, truncated, info

    def stabilities_freezing(self):
        out = super().stabilities_freezing()
        TRACE.append(dict(op="stabilities_freezing", result=[bool(out[0]), bool(out[1])]))
        return out

    _nested = False

    def create_block(self, action):
        b = super().create_block(action)
        if not self._nested:                # (the oracle's collision_on_action poses the block through create_block)
            TRACE.append(dict(op="create_block", action=act(action), block=blk(b)))
        return b

    def collision_on_action(self, action, xlim, ylim):
        self._nested = True
        try:
            out = super().collision_on_action(action, xlim, ylim)
        finally:
            self._nested = False
        TRACE.append(dict(op="collision_on_action", action=act(action), xlim=[fx(v) for v in xlim],
                          ylim=[fx(v) for v in ylim], result=bool(out)))
        return out


def recording_render(blocks, xlim, ylim, img_size=(512, 512)):
    blocks = list(blocks)
    img = orend.render_blocks_2d(blocks, xlim=xlim, ylim=ylim, img_size=img_size)
    TRACE.append(dict(op="render_blocks_2d", blocks=[blk(b) for b in blocks], xlim=[fx(v) for v in xlim],
                      ylim=[fx(v) for v in ylim], img_size=list(img_size), pixels=int(img.sum()), sha1=bits_digest(img)))
    return img


def _not_part_of_the_env(*a, **k):
    raise NotImplementedError("plotting helper: not used by rollout_episode(log_images=False)")


class Frame(tuple):
    """The oracle keeps the face frame of assembly_env.py:118-124 as a (point, normal) pair of the xz-plane and
    unpacks it as such; the reference's callers read compas' Frame attributes.  This pair offers both."""

    @property
    def point(self):
        return [self[0][0], 0.0, self[0][1]]

    @property
    def normal(self):
        return [self[1][0], 0.0, self[1][1]]

    @property
    def xaxis(self):
        return [self[1][1], 0.0, -self[1][0]]

    zaxis = normal


_plain_frame = oae.Shape.get_face_frame_2d
oae.Shape.get_face_frame_2d = lambda self, face: Frame(_plain_frame(self, face))


def install_modules():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m
    mod("assembly_gym")
    mod("assembly_gym.envs")
    mod("assembly_gym.utils")
    mod("assembly_gym.envs.gym_env", AssemblyGym=RecordingGym, Action=ogym.Action, sparse_reward=ogym.sparse_reward,
        tower_setup=ogym.tower_setup, hard_tower_setup=ogym.hard_tower_setup, bridge_setup=ogym.bridge_setup,
        horizontal_bridge_setup=ogym.horizontal_bridge_setup)
    mod("assembly_gym.envs.assembly_env", AssemblyEnv=oae.AssemblyEnv, Block=oae.Block, Shape=oae.Shape)
    mod("assembly_gym.utils.rendering", render_blocks_2d=recording_render, get_rgb_array=_not_part_of_the_env,
        plot_cra_assembly=_not_part_of_the_env, render_assembly_env=_not_part_of_the_env)
    # experiment trackers / plotting: imported at module level by successor_dqn.py, unused on this path
    mod("aim")
    mod("wandb")
    plt = mod("matplotlib.pyplot")
    mod("matplotlib", pyplot=plt)
    sys.path.insert(0, REFERENCE)


def tensor_digest(t):
    return hashlib.sha1(np.ascontiguousarray(t.detach().cpu().numpy()).tobytes()).hexdigest()


def main():
    install_modules()
    from robotoddler.training import successor_dqn as sdqn          # the reference's module, unmodified
    from robotoddler.models.cv import SuccessorMLP
    from robotoddler.utils.replay_memory import ReplayBuffer
    from robotoddler.utils.utils import init_weights
    random.seed(3)
    np.random.seed(3)
    torch.manual_seed(3)
    img_size = (64, 64)
    xlim, ylim = (-3, 7), (0., 10)
    x_discr_ground = np.linspace(-2, 0, 10)                          # successor_dqn.py:611-616
    hidden = [256, 128, 64, 128, 256]                               # successor_dqn.py:626
    policy_net = SuccessorMLP(img_size=img_size, hidden_dims=hidden)
    target_net = SuccessorMLP(img_size=img_size, hidden_dims=hidden)
    policy_net.apply(init_weights)
    target_net.load_state_dict(policy_net.state_dict())
    optimizer = torch.optim.Adam(policy_net.parameters(), lr=0.01)
    replay = ReplayBuffer(capacity=2000)
    eps_greedy = sdqn.EpsilonGreedy(eps_start=0.5, gamma=0.999, eps_end=0.05, episode=0)

    def setup_fct():                                                # successor_dqn.py:688-689 (bridge_length = 1)
        return ogym.horizontal_bridge_setup(num_obstacles=1)

    env = RecordingGym(reward_fct=ogym.sparse_reward, max_steps=10, restrict_2d=True,
                       assembly_env=oae.AssemblyEnv(render=False))  # successor_dqn.py:695
    episodes, losses = [], []
    for ep in range(3):
        TRACE.append(dict(op="episode", index=ep))
        transitions, _ = sdqn.rollout_episode(env, eps_greedy.step(), policy_net, x_discr_ground=x_discr_ground,
                                              setup_fct=setup_fct, offset_values=[0], img_size=img_size, xlim=xlim,
                                              ylim=ylim, log_images=False, device=None)
        replay.push(transitions)
        episodes.append([dict(action=act(t.action), reward=float(t.reward), lin_reward=float(t.lin_reward), done=bool(t.done),
                              n_next_actions=len(t.next_available_actions), td_error=float(t.td_error),
                              block_features=tensor_digest(t.block_features), binary_features=t.binary_features.flatten().tolist(),
                              action_features=tensor_digest(t.action_features),
                              next_block_features=tensor_digest(t.next_block_features[:1]),
                              next_binary_features=t.next_binary_features[0].tolist(),
                              next_actions_features=tensor_digest(t.next_actions_features))
                         for t in transitions])
        out = sdqn.train_policy_net(policy_net, target_net, optimizer, replay, gamma=0.8,
                                    loss_fct="mse_q_values+mse_block_features", n_steps=2, batch_size=4, device="cpu")
        losses.append(out)
        sdqn.update_target_net(policy_net, target_net, tau=0.01)
    doc = dict(about="call trace of the reference's unmodified rollout_episode / train_policy_net against the recording "
                     "oracle view of assembly_gym; generated by tests/golden/make_reference_rollout.py",
               reference_files=["robotoddler/training/successor_dqn.py", "robotoddler/utils/actions.py",
                                "robotoddler/models/cv.py", "robotoddler/utils/replay_memory.py"],
               config=dict(model="SuccessorMLP", hidden_dims=hidden, loss="mse_q_values+mse_block_features", max_steps=10,
                           setup="horizontal_bridge_setup(num_obstacles=1)", seed=3, episodes=3),
               xlim=[fx(v) for v in xlim], ylim=[fx(v) for v in ylim], x_discr_ground=[fx(v) for v in x_discr_ground],
               calls=TRACE, transitions=episodes, train_losses=losses)
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "reference_rollout_trace.json.gz")
    import gzip
    with open(path, "wb") as raw, gzip.GzipFile(fileobj=raw, mode="wb", mtime=0) as fh:
        fh.write(json.dumps(doc, separators=(",", ":")).encode())
    ops = {}
    for c in TRACE:
        ops[c["op"]] = ops.get(c["op"], 0) + 1
    print(path, os.path.getsize(path), "bytes;", ops, "; transitions per episode", [len(e) for e in episodes], "; losses", losses)


if __name__ == "__main__":
    main()
