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
