"""Shared helpers of the parity tests: run the same scripted actions through the CPU oracle
(oracle/, the checker) and through the CUDA path (bridges_b200, via the C ABI)."""
import numpy as np

from oracle import stability as ost
from oracle.assembly_env import AssemblyEnv as OAssemblyEnv
from oracle.assembly_env import Shape as OShape
from oracle.gym_env import Action as OAction
from oracle.gym_env import AssemblyGym as OAssemblyGym
from oracle.gym_env import sparse_reward as o_sparse_reward
from oracle.rendering import render_blocks_2d as o_render

XLIM = (-3.0, 7.0)
YLIM = (0.0, 10.0)
IMG = (64, 64)

URDF = {"trapezoid": "shapes/trapezoid.urdf", "hexagon": "shapes/hexagon.urdf", "cube": "shapes/cube.urdf",
        "cube1": "shapes/cube1.urdf"}


def oracle_env(shape_names, obstacles=(), targets=(), mu=0.8, density=1.0, max_steps=None, shape_kwargs=None):
    shape_kwargs = shape_kwargs or {}
    shapes = [OShape(urdf_file=URDF[n], name=n, **shape_kwargs.get(i, {})) for i, n in enumerate(shape_names)]
    return OAssemblyGym(shapes=shapes, obstacles=list(obstacles), targets=list(targets), reward_fct=o_sparse_reward,
                        restrict_2d=True, max_steps=max_steps, assembly_env=OAssemblyEnv(mu=mu, density=density))


def oracle_trace(env, actions):
    """Per step: everything the CUDA step returns, computed by the oracle with the reference's
    call pattern (step + stabilities_freezing)."""
    trace = []
    for a in actions:
        obs, reward, terminated, truncated, _ = env.step(OAction(*a))
        frozen, unfrozen = env.stabilities_freezing()
        blocks = env.assembly_env.blocks
        trace.append(dict(
            pose=[b.pose for b in blocks],
            raster=o_render(blocks, XLIM, YLIM, IMG),
            stable=bool(obs["stable"]), stable_unfrozen=bool(unfrozen), frozen=bool(frozen),
            reward=reward, terminated=bool(terminated), truncated=bool(truncated) if truncated is not None else False,
            distance=list(obs["distance_to_targets"]), n_reached=len(obs["targets_reached"]),
            n_interfaces=len(env.assembly_env.cra_assembly.interfaces)))
    return trace


def residuals(env):
    """(r_frozen, r_unfrozen) of the oracle's current assembly (BVLS), None where the edge-less rule applies."""
    ae = env.assembly_env
    out = []
    last = len(ae.blocks) - 1
    for release_last in (False, True):
        saved = [b.is_static for b in ae.blocks]
        if release_last:
            ae.blocks[last].is_static = False
        ae._reset_cra_assembly()
        asm = ae.cra_assembly
        if asm.number_of_edges() == 0 or not asm.free_nodes():
            out.append(None)
        else:
            A, b = ost.equilibrium_system(asm, ae.mu, ae.density)
            out.append(ost.equilibrium_residual(A, b, ae.mu))
        for blk, s in zip(ae.blocks, saved):
            blk.is_static = s
        ae._reset_cra_assembly()
    return out


# ---------------------------------------------------------------------------------------------------
# Oracle replay of recorded lock-step trajectories (one job per environment, run in worker processes:
# the oracle is pure Python + HiGHS and the GPU box has 16+ idle host cores).
def bits_of(img):
    """bool [64, 64] -> list of 64 Python ints (bit x of word r = pixel (row r, column x))."""
    w = (1 << np.arange(64, dtype=np.uint64)).astype(np.uint64)
    return [int(v) for v in (img.astype(np.uint64) * w[None, :]).sum(axis=1, dtype=np.uint64)]


def replay_worker(job):
    """job: dict(shapes, obstacles, targets, mu, max_steps, actions=[tuple | None per lock-step iteration],
    x_ground, offsets, cand_steps=set of iterations at which the candidate list / filter mask is wanted too,
    reset_after=iterations after which the environment is reset although its episode did not end by itself).
    The environment is reset after every iteration that ended its episode (terminated | truncated) or offered
    no action (None) -- what `reset_done` does on the GPU.  Returns one dict per iteration."""
    from oracle import actions as oact
    from oracle import features as ofeat
    env = oracle_env(job["shapes"], job["obstacles"], job["targets"], mu=job["mu"], max_steps=job["max_steps"])
    obs, _ = env.reset()
    obstacle_f = o_render(obs["obstacle_blocks"], XLIM, YLIM, IMG).astype(np.float32)[None]
    out = []
    for k, a in enumerate(job["actions"]):
        rec = {}
        if k in job.get("cand_steps", ()):
            block_f, _ = ofeat.get_state_features(obs, XLIM, YLIM, IMG)
            cands = [*oact.generate_actions(env, job["x_ground"], list(job["offsets"]))]
            cand_f = ofeat.get_action_features(env, cands, XLIM, YLIM, IMG)
            _, _, mask = oact.filter_actions(env, cands, cand_f, block_f, obstacle_f, XLIM, YLIM)
            rec["cands"] = [(c.target_block, c.target_face, c.shape, c.face, c.offset_x, c.offset_y) for c in cands]
            rec["cand_mask"] = [bool(m) for m in mask]
        if a is None:
            rec["skipped"] = True
            out.append(rec)
            obs, _ = env.reset()
            continue
        obs, reward, terminated, truncated, _ = env.step(OAction(*a))
        frozen, unfrozen = env.stabilities_freezing()
        blocks = env.assembly_env.blocks
        r_frozen, r_unfrozen = residuals(env)
        rec.update(stable=bool(obs["stable"]), stable_unfrozen=bool(unfrozen), frozen=bool(frozen),
                   r_frozen=r_frozen, r_unfrozen=r_unfrozen, reward=float(reward), terminated=bool(terminated),
                   truncated=bool(truncated) if truncated is not None else False, n_blocks=len(blocks),
                   n_interfaces=len(env.assembly_env.cra_assembly.interfaces),
                   distance=[float(d) for d in obs["distance_to_targets"]], n_reached=len(obs["targets_reached"]),
                   bits=bits_of(o_render(blocks, XLIM, YLIM, IMG)), new_bits=bits_of(o_render(blocks[-1:], XLIM, YLIM, IMG)),
                   pose=blocks[-1].pose)
        out.append(rec)
        if terminated or truncated or k in job.get("reset_after", ()):
            obs, _ = env.reset()
    return out


def replay_parallel(jobs, workers=None):
    import multiprocessing as mp
    import os
    workers = workers or min(len(jobs), os.cpu_count() or 1)
    if workers <= 1:
        return [replay_worker(j) for j in jobs]
    with mp.get_context("spawn").Pool(workers) as pool:
        return pool.map(replay_worker, jobs, chunksize=max(1, len(jobs) // (4 * workers)))
