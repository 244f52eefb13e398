#!/usr/bin/env python
"""Successor-feature Q-learning (BASELINE.json configs[1] / configs[2]) on the batched B200 environment: greedy
success rate against wall-clock time, to be read beside the reference's own learner on the restated CPU env
(`tools/reference_learner_cpu.py` -> profiles/r2_reference_learner_cpu_h4.jsonl).

The learner restates what robotoddler/training/successor_dqn.py does, in lock-step form:

  * network: the reference's `SuccessorMLP` architecture (models/cv.py:76-105: one MLP over the four 64 x 64 planes
    block / action / reward / obstacle + 6 binary features, hidden 256-128-64-128-256, outputs a two-channel successor
    image and 2 x 6 binary successor features; q = sum(softmax(successor image)[1] * reward plane));
  * losses `mse_q_values+mse_block_features` (successor_dqn.py:216-230): q(s, a) against lin_reward + gamma * max_a' q_target,
    successor image channel 0 against action plane + gamma * target successor image of the arg-max next candidate;
    zero beyond the end of an episode; soft target update with tau (successor_dqn.py:280-288);
  * epsilon-greedy over the valid candidates (uniform exploration instead of the reference's join-score heuristic).

What differs is the data flow the GPU allows: E environments advance per iteration through the fused rollout
(`bw_rollout_begin` / `bw_rollout_commit`), ONE batched network pass scores the valid candidates of all of them, and the
E fresh transitions of the iteration -- whose next-state candidates are the candidate buffers the rollout has just
refreshed -- are the training batch (no replay memory is needed to decorrelate E parallel episodes).

    python examples/train_successor.py --tower-height 4 --max-steps 15 --envs 256 --seconds 120
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

from bench import X_GROUND, task_def                                             # noqa: E402
from bridges_b200.envs.batched import BatchedAssemblyGym                         # noqa: E402
from bridges_b200.rollout import FusedRollout, record_column                     # noqa: E402

IMG = 64


class SuccessorMLP(nn.Module):
    """The architecture of models/cv.py:76-105 (restated): forward(block, binary, action, reward, obstacle) ->
    (q [R], successor image [R,2,64,64], successor binary [R,2,6])."""

    def __init__(self, hidden=(256, 128, 64, 128, 256), n_binary=6):
        super().__init__()
        dims = [4 * IMG * IMG + n_binary, *hidden]
        layers = []
        for a, b in zip(dims[:-1], dims[1:]):
            layers += [nn.Linear(a, b), nn.ReLU()]
        layers.append(nn.Linear(dims[-1], 2 * IMG * IMG + 2 * n_binary))
        self.mlp = nn.Sequential(*layers)
        self.n_binary = n_binary
        for m in self.modules():                       # init_weights, robotoddler/utils/utils.py:12-19
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                m.bias.data.fill_(0.01)

    def forward(self, block, binary, action, reward, obstacle):
        x = torch.cat([block, action, reward, obstacle], dim=1).flatten(1)
        x = self.mlp(torch.cat([x, binary], dim=1))
        succ = x[:, :2 * IMG * IMG].view(-1, 2, IMG, IMG)
        succ_bin = x[:, 2 * IMG * IMG:].view(-1, 2, self.n_binary)
        q = (succ.softmax(dim=1)[:, 1] * reward.squeeze(1)).sum(dim=(-1, -2))
        return q, succ, succ_bin


def score_candidates(net, env, cand, state, reward_f, obstacle_f, chunk_rows, want_best_successor=False):
    """One batched pass over the valid candidates of all environments: q [E, amax] (-inf where invalid), the validity
    mask and, on request, the successor image (channel 0) of every environment's arg-max candidate [E, 64, 64]
    (a second pass over those E rows only)."""
    E, dev, amax = env.num_envs, env.device, cand["amax"]
    valid = cand["valid"].bool() & (torch.arange(amax, device=dev)[None, :] < cand["n"][:, None])
    e_idx, a_idx = valid.nonzero(as_tuple=True)
    q_full = torch.full((E, amax), float("-inf"), device=dev)
    for lo in range(0, e_idx.numel(), chunk_rows):
        er, ar = e_idx[lo:lo + chunk_rows], a_idx[lo:lo + chunk_rows]
        action_f = env.expand_bits(cand["bits"][er, ar].contiguous())
        q, _, _ = net(state["block"][er], state["binary"][er], action_f, reward_f[er], obstacle_f[er])
        q_full[er, ar] = q
    best_s = None
    if want_best_successor:
        best = q_full.argmax(dim=1)
        action_f = env.expand_bits(cand["bits"][torch.arange(E, device=dev), best].contiguous())
        _, succ, _ = net(state["block"], state["binary"], action_f, reward_f, obstacle_f)
        best_s = succ[:, 0]
    return q_full, valid, best_s


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tower-height", type=int, default=4)
    ap.add_argument("--max-steps", type=int, default=15)
    ap.add_argument("--envs", type=int, default=256)
    ap.add_argument("--seconds", type=float, default=120.0, help="wall-clock budget")
    ap.add_argument("--gamma", type=float, default=0.8)            # successor_dqn.py defaults
    ap.add_argument("--lr", type=float, default=1e-3)
    ap.add_argument("--tau", type=float, default=0.01)
    ap.add_argument("--eval-every", type=int, default=25, help="iterations between greedy evaluations")
    ap.add_argument("--chunk-rows", type=int, default=16384)
    ap.add_argument("--log", default=os.path.join(ROOT, "gpurun_out", "train_successor.jsonl"))
    args = ap.parse_args()
    torch.manual_seed(0)
    E = args.envs
    amax = 128 if args.max_steps <= 10 else 256
    env = BatchedAssemblyGym(E, ["shapes/trapezoid.urdf"], max_steps=args.max_steps)
    env.reset(task_def(args.tower_height))
    ev_env = BatchedAssemblyGym(E, ["shapes/trapezoid.urdf"], max_steps=args.max_steps)     # greedy evaluation
    dev = env.device
    feats = env.observe(block=False, binary=False, obstacle=True, reward=True)
    reward_f, obstacle_f = feats["reward"], feats["obstacle"]
    roll = FusedRollout(env, X_GROUND, (0.0,), amax=amax, chunk_steps=1)
    policy_net, target_net = SuccessorMLP().to(dev), SuccessorMLP().to(dev)
    target_net.load_state_dict(policy_net.state_dict())
    opt = torch.optim.Adam(policy_net.parameters(), lr=args.lr)
    gen = torch.Generator(device=dev)
    gen.manual_seed(0)
    os.makedirs(os.path.dirname(args.log), exist_ok=True)
    log = open(args.log, "w")
    t0 = time.perf_counter()
    it, env_steps, episodes, successes = 0, 0, 0, 0

    def evaluate():
        ev_env.reset(task_def(args.tower_height))
        ev_roll = FusedRollout(ev_env, X_GROUND, (0.0,), amax=amax, chunk_steps=1)
        done_once = torch.zeros(E, dtype=torch.bool, device=dev)
        ok = torch.zeros(E, dtype=torch.bool, device=dev)
        with torch.no_grad():
            for _ in range(args.max_steps + 1):
                cand = ev_roll.candidates()
                state = ev_env.observe(block=True, binary=True)
                q, valid, _ = score_candidates(policy_net, ev_env, cand, state, reward_f, obstacle_f, args.chunk_rows)
                index = torch.where(valid.any(dim=1), q.argmax(dim=1), torch.full((E,), -1, device=dev)).to(torch.int32)
                rec = ev_roll.step(index)
                done = (record_column(rec, "done") != 0) & (record_column(rec, "valid") != 0)
                first = done & ~done_once
                ok |= first & (record_column(rec, "reward") >= 1.0)
                done_once |= done
        return float(ok.float().mean()), float(done_once.float().mean())

    while time.perf_counter() - t0 < args.seconds:
        eps = max(0.05, 0.5 * (0.995 ** it))
        cand = roll.candidates()
        state = env.observe(block=True, binary=True)
        with torch.no_grad():
            q, valid, _ = score_candidates(policy_net, env, cand, state, reward_f, obstacle_f, args.chunk_rows)
        greedy = q.argmax(dim=1)
        noise = torch.rand((E, amax), device=dev, generator=gen).masked_fill(~valid, -1.0)
        explore = torch.rand(E, device=dev, generator=gen) < eps
        index = torch.where(explore, noise.argmax(dim=1), greedy)
        index = torch.where(valid.any(dim=1), index, torch.full_like(index, -1)).to(torch.int32)
        chosen_bits = cand["bits"][torch.arange(E, device=dev), index.clamp(min=0).long()].clone()
        rec = roll.step(index).clone()
        has = record_column(rec, "valid") != 0
        done = record_column(rec, "done") != 0
        lin_reward, reward = record_column(rec, "lin_reward"), record_column(rec, "reward")
        # TD targets from the candidates of the next states (the buffers the rollout has just refreshed)
        with torch.no_grad():
            nxt_cand = roll.candidates()
            nxt_state = env.observe(block=True, binary=True)
            nq, nvalid, nsucc = score_candidates(target_net, env, nxt_cand, nxt_state, reward_f, obstacle_f, args.chunk_rows, True)
            next_q = torch.where(nvalid.any(dim=1), nq.max(dim=1).values, torch.zeros(E, device=dev))
            alive = (~done).float()
            action_f = env.expand_bits(chosen_bits)
            y_q = lin_reward + args.gamma * next_q * alive
            y_s = action_f[:, 0] + args.gamma * nsucc * alive[:, None, None]
        sel = has.nonzero(as_tuple=False).flatten()
        if sel.numel() > 0:
            policy_net.train()
            qv, succ, _ = policy_net(state["block"][sel], state["binary"][sel], action_f[sel], reward_f[sel], obstacle_f[sel])
            loss = torch.nn.functional.mse_loss(qv, y_q[sel]) + torch.nn.functional.mse_loss(succ[:, 0], y_s[sel])
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
            with torch.no_grad():                        # update_target_net, successor_dqn.py:280-288
                for p, tp in zip(policy_net.parameters(), target_net.parameters()):
                    tp.mul_(1.0 - args.tau).add_(p, alpha=args.tau)
        it += 1
        env_steps += int(has.sum())
        episodes += int((done & has).sum())
        successes += int((done & has & (reward >= 1.0)).sum())
        if it % args.eval_every == 0:
            g_succ, g_done = evaluate()
            row = dict(iteration=it, wall_s=time.perf_counter() - t0, env_steps=env_steps, episodes=episodes,
                       explore_success=successes / max(episodes, 1), greedy_success=g_succ, greedy_finished=g_done,
                       epsilon=eps, loss=float(loss.detach()) if sel.numel() else None, tower_height=args.tower_height,
                       max_steps=args.max_steps, envs=E, learner="SuccessorMLP (models/cv.py:76-105 restated), "
                       "mse_q_values+mse_block_features, online batches of E fresh transitions")
            print(json.dumps(row), flush=True)
            log.write(json.dumps(row) + "\n")
            log.flush()
            episodes, successes = 0, 0
    env.close()
    ev_env.close()


if __name__ == "__main__":
    main()
