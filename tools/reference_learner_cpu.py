"""The reference's own learner on the restated CPU env: success rate against wall-clock (BASELINE.json configs[2]).

TEST / MEASUREMENT INFRASTRUCTURE, build container only (needs /root/reference).  The reference's UNMODIFIED
`rollout_episode`, `EpsilonGreedy`, `train_policy_net`, `update_target_net` (robotoddler/training/successor_dqn.py),
`SuccessorMLP` (models/cv.py:76-105) and `ReplayBuffer` (utils/replay_memory.py) run the main loop of
successor_dqn.py:698-760 (`--model=SuccessorMLP --loss_function=mse_q_values+mse_block_features --max_steps=15`,
defaults otherwise) on the `tower_height=k` task, with `assembly_gym` bound to this repository's CPU oracle (the real
package cannot be installed here).  Every `--evaluate-every` episodes a greedy rollout is run, as the script does.
Output: one JSON line per evaluation (wall-clock seconds, episodes, env steps, greedy success, mean training reward).

    python tools/reference_learner_cpu.py --tower-height 4 --episodes 600 --out profiles/r2_reference_learner_cpu_h4.jsonl
"""
import argparse
import json
import os
import random
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tower-height", type=int, default=4)
    ap.add_argument("--max-steps", type=int, default=15)
    ap.add_argument("--episodes", type=int, default=600)
    ap.add_argument("--evaluate-every", type=int, default=25)
    ap.add_argument("--eval-episodes", type=int, default=1)
    ap.add_argument("--threads", type=int, default=4)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r2_reference_learner_cpu.jsonl"))
    args = ap.parse_args()
    torch.set_num_threads(args.threads)
    import make_reference_rollout as rec                     # binds assembly_gym -> oracle, stubs aim / wandb / matplotlib
    rec.install_modules()
    from oracle import assembly_env as oae
    from oracle import gym_env as ogym
    sys.modules["assembly_gym.envs.gym_env"].AssemblyGym = ogym.AssemblyGym       # plain oracle env, no call recording
    sys.modules["assembly_gym.utils.rendering"].render_blocks_2d = __import__("oracle.rendering", fromlist=["x"]).render_blocks_2d
    from robotoddler.training import successor_dqn as sdqn
    from robotoddler.models.cv import SuccessorMLP
    from robotoddler.utils.replay_memory import ReplayBuffer
    from robotoddler.utils.utils import init_weights
    random.seed(args.seed); np.random.seed(args.seed); torch.manual_seed(args.seed)
    img_size, xlim, ylim = (64, 64), (-3, 7), (0., 10)
    x_discr_ground = np.linspace(-2, 0, 10)
    hidden = [256, 128, 64, 128, 256]
    policy_net = SuccessorMLP(img_size=img_size, hidden_dims=hidden)
    target_net = SuccessorMLP(img_size=img_size, hidden_dims=hidden)
    policy_net.apply(init_weights)
    target_net.load_state_dict(policy_net.state_dict())
    optimizer = torch.optim.Adam(policy_net.parameters(), lr=0.01)           # --learning_rate default
    replay = ReplayBuffer(capacity=2000)                                     # --replay_buffer_capacity default
    eps_greedy = sdqn.EpsilonGreedy(eps_start=0.5, gamma=0.999, eps_end=0.05, episode=0, max_steps=args.max_steps)
    greedy = lambda q, *a, **k: torch.argmax(q)

    def setup_fct():
        return ogym.tower_height_setup(args.tower_height)

    env = ogym.AssemblyGym(reward_fct=ogym.sparse_reward, max_steps=args.max_steps, restrict_2d=True,
                           assembly_env=oae.AssemblyEnv(render=False))
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    log = open(args.out, "w")
    t0 = time.perf_counter()
    env_steps, t_env, rewards, succ = 0, 0.0, [], []
    for ep in range(1, args.episodes + 1):
        ta = time.perf_counter()
        transitions, _ = sdqn.rollout_episode(env, eps_greedy.step(), policy_net, x_discr_ground=x_discr_ground,
                                              setup_fct=setup_fct, offset_values=[0], img_size=img_size, xlim=xlim, ylim=ylim,
                                              log_images=False, device=None)
        t_env += time.perf_counter() - ta
        env_steps += len(transitions)
        rewards.append(float(transitions[-1].reward))
        succ.append(float(transitions[-1].reward) >= 1.0)
        replay.push(transitions)
        sdqn.train_policy_net(policy_net, target_net, optimizer, replay, gamma=0.8, loss_fct="mse_q_values+mse_block_features",
                              n_steps=20, batch_size=32, device="cpu")        # script defaults
        sdqn.update_target_net(policy_net, target_net, tau=0.01)
        if ep % args.evaluate_every == 0:
            ok = 0
            for _ in range(args.eval_episodes):
                tr, _ = sdqn.rollout_episode(env, greedy, policy_net, x_discr_ground=x_discr_ground, setup_fct=setup_fct,
                                             offset_values=[0], img_size=img_size, xlim=xlim, ylim=ylim, log_images=False, device=None)
                ok += float(tr[-1].reward) >= 1.0
            row = dict(episode=ep, wall_s=time.perf_counter() - t0, rollout_s=t_env, env_steps=env_steps,
                       greedy_success=ok / args.eval_episodes, explore_success_last=float(np.mean(succ[-args.evaluate_every:])),
                       mean_final_reward_last=float(np.mean(rewards[-args.evaluate_every:])), epsilon=eps_greedy.epsilon,
                       tower_height=args.tower_height, max_steps=args.max_steps, threads=args.threads,
                       learner="reference successor_dqn.py (unmodified), SuccessorMLP, mse_q_values+mse_block_features",
                       env="restated reference CPU env (oracle/)")
            log.write(json.dumps(row) + "\n")
            log.flush()
            print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
