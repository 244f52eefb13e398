"""Scratch: harvest whole episodes of bench-like rollouts from the CPU oracle with the identity of every contact
column, for offline experiments with warm-started verdict solvers (tools/simplex_lab.py).  Test infrastructure only.

Per env step one record: the equilibrium system with EVERY block released (rows of blocks 0..n-1; the frozen
problem of the same step is that system without the last block's three rows), the interface list as (a, b) node
pairs (a = -1: floor), the oracle's two verdicts.

python tools/harvest_episodes.py OUT.pkl [seconds_per_worker] [task: tower2|tower4|bridge]"""
import os, sys, time, pickle
import multiprocessing as mp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def worker(job):
    seed, budget_s, task = job
    import numpy as np
    from bench import task_def, bridge_def, X_GROUND
    from oracle import actions as oact
    from oracle import features as ofeat
    from oracle import stability as ost
    from oracle.assembly_env import AssemblyEnv, Shape
    from oracle.gym_env import AssemblyGym, sparse_reward
    from oracle.rendering import render_blocks_2d
    rng = np.random.default_rng(seed)
    xlim, ylim, img = (-3.0, 7.0), (0.0, 10.0), (64, 64)
    if task == "bridge":
        t, names, max_steps = bridge_def(5), ["trapezoid", "hexagon"], 15
    elif task == "tower4":
        t, names, max_steps = task_def(4), ["trapezoid"], 15
    else:
        t, names, max_steps = task_def(2), ["trapezoid"], 10
    env = AssemblyGym(shapes=[Shape(urdf_file=f"shapes/{nm}.urdf", name=nm) for nm in names], obstacles=t["obstacles"],
                      targets=t["targets"], reward_fct=sparse_reward, restrict_2d=True, max_steps=max_steps,
                      assembly_env=AssemblyEnv())
    out = []
    episode = 0
    t_end = time.perf_counter() + budget_s
    while time.perf_counter() < t_end:
        obs, _ = env.reset()
        episode += 1
        obstacle_f = render_blocks_2d(obs['obstacle_blocks'], xlim, ylim, img).astype(np.float32)[None]
        done = False
        while not done and time.perf_counter() < t_end:
            block_f, _ = ofeat.get_state_features(obs, xlim, ylim, img)
            cands = [*oact.generate_actions(env, X_GROUND, [0.0])]
            cand_f = ofeat.get_action_features(env, cands, xlim, ylim, img)
            kept, _, _ = oact.filter_actions(env, cands, cand_f, block_f, obstacle_f, xlim, ylim)
            if not kept:
                break
            action = kept[int(rng.integers(len(kept)))]
            obs, reward, terminated, truncated, _ = env.step(action)
            ae = env.assembly_env
            n = len(ae.blocks)
            frozen_ok, released_ok = env.stabilities_freezing()
            ae.unfreeze_block(n - 1)
            asm = ae.cra_assembly
            rec = {"episode": (seed, episode), "n_blocks": n, "mu": ae.mu, "frozen_ok": frozen_ok, "released_ok": released_ok,
                   "free": list(asm.free_nodes()), "itf": [(it.a, it.b) for it in asm.interfaces]}
            if asm.number_of_edges() == 0:
                rec["A"], rec["b"] = None, None
            else:
                rec["A"], rec["b"] = ost.equilibrium_system(asm, ae.mu, ae.density)
            ae.freeze_block(n - 1)
            out.append(rec)
            done = bool(terminated or truncated)
    return out


if __name__ == "__main__":
    path = sys.argv[1]
    budget = float(sys.argv[2]) if len(sys.argv) > 2 else 60.0
    task = sys.argv[3] if len(sys.argv) > 3 else "bridge"
    n = os.cpu_count()
    with mp.get_context("spawn").Pool(n) as pool:
        res = pool.map(worker, [(9000 + i, budget, task) for i in range(n)])
    recs = [r for w in res for r in w]
    with open(path, "wb") as fh:
        pickle.dump(recs, fh)
    import collections
    print(len(recs), "steps", sorted(collections.Counter(r["n_blocks"] for r in recs).items()))
