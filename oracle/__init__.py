"""CPU oracle for the assembly_gym env step -- TEST INFRASTRUCTURE ONLY.

This package is a float64 CPU restatement of the reference's hot path
(`assembly_gym/assembly_gym/envs/{assembly_env,gym_env}.py`,
`assembly_gym/assembly_gym/utils/{geometry,stability,rendering}.py`,
`robotoddler/utils/actions.py`, the feature functions of
`robotoddler/training/successor_dqn.py`) plus the published algorithms of the
un-vendored third-party packages that path calls (compas 2.1.1 mesh/geometry
helpers, compas_cra `assembly_interfaces_numpy` / `rbe_solve`).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` leg may import it, and only as the checker / the timed CPU
baseline.  The product (`bridges_b200`) never imports it and has no CPU
fallback.

Parity pin status:
  * block library face tables, placement geometry, target bookkeeping, reward,
    distances: pinned by the reference's stored notebook outputs
    (tests/golden/notebook_goldens.json) and by direct comparison with the
    reference's shape files;
  * stable/unstable verdicts: pinned by the expected labels of
    `assembly_gym/utils/structures.py:22-108` and the notebook runs;
  * contact-force VALUES, interface tolerances and the bit pattern of rotated
    poses: PARITY UNPINNED (the arithmetic lives in compas / compas_cra /
    IPOPT, none of which is present under /root/reference nor installable
    here).  The oracle defines a canonical arithmetic for those and the CUDA
    path is proven equal to it.
"""
