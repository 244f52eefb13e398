#!/usr/bin/env python
"""Demonstration: a Q-network trained against the batched B200 environment on the `tower_height=k`
task (BASELINE.json configs[1] / configs[2]), success rate against wall-clock time.

This is NOT robotoddler's learner (out of scope): it only shows the
pieces of this repository working together the way `successor_dqn.py` would use them -- lock-step
environments, the candidate kernel, one batched Q-network pass per step (`rollout.q_network_policy`,
the network has the reference's 5-argument signature, models/cv.py:76-105), the fused rollout
(`bw_rollout_begin` / `bw_rollout_commit`) with its packed transition records on the device -- with the simplest
possible learning rule (regression of Q(s, a) on Monte-Carlo returns).

    python examples/train_tower.py --tower-height 2 --envs 1024 --iters 30
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bench import X_GROUND, task_def                                  # noqa: E402
from bridges_b200.envs.batched import BatchedAssemblyGym              # noqa: E402
from bridges_b200.rollout import FusedRollout, q_network_policy, record_column   # noqa: E402


class QNet(nn.Module):
    """forward(block, binary, action, reward, obstacle) -> (q, None, None), as the reference's networks."""

    def __init__(self):
        super().__init__()
        self.conv = nn.Sequential(nn.Conv2d(4, 16, 5, stride=2, padding=2), nn.ReLU(),
                                  nn.Conv2d(16, 32, 3, stride=2, padding=1), nn.ReLU(),
                                  nn.Conv2d(32, 64, 3, stride=2, padding=1), nn.ReLU(),
                                  nn.AdaptiveAvgPool2d(4), nn.Flatten())
        self.head = nn.Sequential(nn.Linear(64 * 16 + 6, 256), nn.ReLU(), nn.Linear(256, 1))

    def forward(self, block, binary, action, reward, obstacle):
        x = self.conv(torch.cat([block, action, reward * 50.0, obstacle], dim=1))
        return self.head(torch.cat([x, binary], dim=1)).squeeze(1), None, None


def collect(roll, policy, steps, gamma):
    """`steps` lock-step iterations of the fused rollout around `policy`; returns the packed records of the
    transitions whose episode ended inside the chunk with their Monte-Carlo returns."""
    E, dev = roll.env.num_envs, roll.env.device
    recs = []
    for _ in range(steps):
        recs.append(roll.step(policy(roll.env, roll.candidates())).clone())
    R = torch.stack(recs)                                             # [T, E, REC] uint8
    reward, done = record_column(R, "reward"), record_column(R, "done") != 0
    has = record_column(R, "valid") != 0
    ret = torch.zeros_like(reward)
    complete = torch.zeros_like(done)
    g = torch.zeros(E, device=dev)
    c = torch.zeros(E, dtype=torch.bool, device=dev)
    for t in range(steps - 1, -1, -1):
        g = torch.where(done[t], reward[t], reward[t] + gamma * g)
        c = done[t] | c
        ret[t], complete[t] = g, c
    keep = (complete & has).reshape(-1)
    flat = R.reshape(steps * E, -1)[keep]
    binary = record_column(flat, "binary")
    episodes = int((done & has).sum())
    success = int((done & has & (reward >= 1.0)).sum())
    return dict(block=record_column(flat, "block_bits"), action=record_column(flat, "action_bits"),
                binary=((binary[:, None] >> torch.arange(6, device=dev)) & 1).float(),
                ret=ret.reshape(-1)[keep]), episodes, success


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tower-height", type=int, default=2)
    ap.add_argument("--envs", type=int, default=1024)
    ap.add_argument("--max-steps", type=int, default=10)
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--chunk", type=int, default=20, help="lock-step env steps collected per iteration")
    ap.add_argument("--train-steps", type=int, default=40)
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--gamma", type=float, default=0.95)
    ap.add_argument("--log", default=os.path.join(ROOT, "gpurun_out", "train_tower.jsonl"))
    args = ap.parse_args()
    torch.manual_seed(0)
    env = BatchedAssemblyGym(args.envs, ["shapes/trapezoid.urdf"], max_steps=args.max_steps)
    env.reset(task_def(args.tower_height))
    dev = env.device
    feats = env.observe(block=False, binary=False, obstacle=True, reward=True)
    roll = FusedRollout(env, X_GROUND, (0.0,), amax=128 if args.max_steps <= 10 else 256, chunk_steps=1)
    net = QNet().to(dev)
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    buf = None
    os.makedirs(os.path.dirname(args.log), exist_ok=True)
    log = open(args.log, "w")
    t0 = time.perf_counter()
    env_steps = 0
    for it in range(args.iters):
        eps = max(0.05, 0.9 * (0.85 ** it))
        policy = q_network_policy(net, feats["reward"], feats["obstacle"], epsilon=eps, seed=it)
        data, episodes, success = collect(roll, policy, args.chunk, args.gamma)
        env_steps += args.chunk * args.envs
        buf = data if buf is None else {k: torch.cat([buf[k], data[k]])[-400000:] for k in data}
        n = buf["ret"].numel()
        loss_v = 0.0
        net.train()
        for _ in range(args.train_steps):
            idx = torch.randint(0, n, (args.batch,), device=dev)
            block = env.expand_bits(buf["block"][idx].contiguous())
            action = env.expand_bits(buf["action"][idx].contiguous())
            q, _, _ = net(block, buf["binary"][idx], action, feats["reward"][:1].expand(args.batch, -1, -1, -1),
                          feats["obstacle"][:1].expand(args.batch, -1, -1, -1))
            loss = torch.nn.functional.smooth_l1_loss(q, buf["ret"][idx])
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
            loss_v = float(loss.detach())
        # greedy evaluation
        env.reset(task_def(args.tower_height))
        roll.begin()
        _, ev_episodes, ev_success = collect(roll, q_network_policy(net, feats["reward"], feats["obstacle"]),
                                             args.max_steps + 2, args.gamma)
        row = dict(iter=it, wall_s=time.perf_counter() - t0, env_steps=env_steps, epsilon=eps, loss=loss_v,
                   explore_success=success / max(episodes, 1), greedy_success=ev_success / max(ev_episodes, 1),
                   greedy_episodes=ev_episodes, buffer=n, tower_height=args.tower_height, envs=args.envs)
        print(json.dumps(row), flush=True)
        log.write(json.dumps(row) + "\n")
        log.flush()
    env.close()


if __name__ == "__main__":
    main()
