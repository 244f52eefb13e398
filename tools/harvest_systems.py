"""Scratch: harvest the equilibrium systems (A, b, mu) of bench-like rollouts from the CPU oracle, for offline
experiments with the solver's iteration strategy (tools/solver_lab.py).  Test infrastructure only.

python tools/harvest_systems.py OUT.pkl [seconds_per_worker] [tower_height] [max_steps]"""
import os, sys, time, pickle
import multiprocessing as mp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def worker(job):
    seed, budget_s, tower_height, max_steps = job
    import numpy as np
    from bench import task_def, X_GROUND
    from oracle import actions as oact
    from oracle import features as ofeat
    from oracle import stability as ost
    from oracle.assembly_env import AssemblyEnv, Shape
    from oracle.gym_env import AssemblyGym, sparse_reward
    from oracle.rendering import render_blocks_2d
    rng = np.random.default_rng(seed)
    xlim, ylim, img = (-3.0, 7.0), (0.0, 10.0), (64, 64)
    t = task_def(tower_height)
    env = AssemblyGym(shapes=[Shape(urdf_file="shapes/trapezoid.urdf", name="trapezoid")], obstacles=t["obstacles"],
                      targets=t["targets"], reward_fct=sparse_reward, restrict_2d=True, max_steps=max_steps,
                      assembly_env=AssemblyEnv())
    out = []
    t_end = time.perf_counter() + budget_s
    while time.perf_counter() < t_end:
        obs, _ = env.reset()
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
            rec = {"n_blocks": len(ae.blocks), "mu": ae.mu}
            for tag in ("frozen", "unfrozen"):
                if tag == "unfrozen":
                    ae.unfreeze_block(len(ae.blocks) - 1)
                asm = ae.cra_assembly
                if asm.number_of_edges() == 0:
                    rec[tag] = None
                else:
                    A, b = ost.equilibrium_system(asm, ae.mu, ae.density)
                    free = asm.free_nodes()
                    coms = [asm.bodies[n + 1].com for n in free]
                    L0 = max([body.radius for body in asm.bodies] + [1e-300])
                    rec[tag] = (A, b, ost.rbe_feasible(A, b, ae.mu))
                    rec[tag + "_geo"] = (free, coms, L0)
                if tag == "unfrozen":
                    ae.freeze_block(len(ae.blocks) - 1)
            out.append(rec)
            done = bool(terminated or truncated)
    return out


if __name__ == "__main__":
    path = sys.argv[1]
    budget = float(sys.argv[2]) if len(sys.argv) > 2 else 60.0
    th = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    ms = int(sys.argv[4]) if len(sys.argv) > 4 else 10
    n = os.cpu_count()
    with mp.get_context("spawn").Pool(n) as pool:
        res = pool.map(worker, [(5000 + i, budget, th, ms) for i in range(n)])
    recs = [r for w in res for r in w]
    with open(path, "wb") as fh:
        pickle.dump(recs, fh)
    import collections
    print(len(recs), "steps", collections.Counter(r["n_blocks"] for r in recs))
