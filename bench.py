#!/usr/bin/env python
"""bench.py -- stability-checked env steps/sec of the batched assembly_gym step on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W   # restated reference CPU env

Workload (BASELINE.json configs[1]): 1024 lock-step envs per GPU, the `tower_height=2` task
(one 0.6-cube obstacle at x=0.6, target just above it; trapezoid blocks; max_steps=10),
synthetic uniformly-random valid actions from the candidate kernel, auto-reset on termination.
One "step" = one `bw_step` over the whole batch: block placement, contact interfaces, the two
equilibrium verdicts (new block frozen / released), targets, reward, termination, raster
update and the f32 [E,1,64,64] observation + [E,6] binary features.

Timing: every step is bracketed by CUDA events on the stream the kernel is launched on; the L2
is flushed (256 MiB write) between steps, outside the event pairs; the whole loop is bracketed by
barrier + synchronize and the max over ranks is reported.  Rank 0 prints ONE JSON line.
"""
import argparse
import ctypes as C
import json
import multiprocessing as mp
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "stability-checked env steps/sec"
UNIT = "env_steps/s"
X_GROUND = [-2.0 + 2.0 * i / 9 for i in range(10)]   # np.linspace(-2, 0, 10), successor_dqn.py:611


def task_def(tower_height, square=0.6):
    obstacles = [(square, 0, i * square + square / 2) for i in range(tower_height - 1)]
    targets = [(square, 0, (tower_height - 1) * square + square / 2)]
    return dict(obstacles=obstacles, targets=targets)


def bridge_def(num_obstacles, square=0.6):
    """horizontal_bridge_setup, gym_env.py:24-42: a row of obstacle cubes on the floor, the target beyond it"""
    obstacles = [(i * square, 0, square / 2) for i in range(1, num_obstacles + 1)]
    targets = [(num_obstacles * square + 2.5 * square, 0, square / 2)]
    return dict(obstacles=obstacles, targets=targets)


def task_spec(args):
    """Picklable description of the benchmark task: obstacles, targets and the block library."""
    if args.task == "bridge":
        t = bridge_def(args.num_obstacles)
        shapes = [x.strip() for x in args.shapes.split(",") if x.strip()]
    else:
        t = task_def(args.tower_height)
        shapes = ["trapezoid"]
    return dict(kind=args.task, shapes=shapes, obstacles=t["obstacles"], targets=t["targets"])


def config_dict(args, n_gpus):
    if args.task == "bridge":
        return {"workload": f"horizontal_bridge_setup(num_obstacles={args.num_obstacles}) with shapes [{args.shapes}], "
                            f"{args.envs} lock-step envs per GPU, max_steps={args.max_steps}, random valid actions, "
                            f"auto-reset (BASELINE.json configs[4] shape)",
                "envs_per_gpu": args.envs, "num_obstacles": args.num_obstacles, "shapes": args.shapes,
                "max_steps": args.max_steps, "image": "64x64 f32", "mu": 0.8,
                "parallelism": f"env-sharded x{n_gpus}, no collective on the step path",
                "cache": "L2 flushed between timed steps (256 MiB write outside the event pairs)"}
    return {"workload": f"tower_height={args.tower_height} trapezoid task, {args.envs} lock-step envs per GPU, "
                        f"max_steps={args.max_steps}, random valid actions, auto-reset (BASELINE.json configs[1])",
            "envs_per_gpu": args.envs, "tower_height": args.tower_height, "max_steps": args.max_steps,
            "image": "64x64 f32", "mu": 0.8, "parallelism": f"env-sharded x{n_gpus}, no collective on the step path",
            "cache": "L2 flushed between timed steps (256 MiB write outside the event pairs)"}


# ------------------------------------------------------------------------------ CPU arm
def _cpu_worker(job):
    """Restated reference CPU env (oracle/) with the reference's call pattern per step:
    env.step (placement, 2 interface rebuilds, 2 RBE solves) + stabilities_freezing (1 rebuild,
    3 solves) + get_state_features (1 raster).  Candidate generation for the random policy runs
    outside the timed sections."""
    seed, budget_s, tower_height, max_steps, max_env_steps = job[:5]
    warm_env_steps, min_s = (job[5], job[6]) if len(job) > 5 else (0, 0.0)
    import numpy as np
    from oracle import actions as oact
    from oracle import features as ofeat
    from oracle.assembly_env import AssemblyEnv, Shape
    from oracle.gym_env import AssemblyGym, sparse_reward
    from oracle.rendering import render_blocks_2d
    rng = np.random.default_rng(seed)
    xlim, ylim, img = (-3.0, 7.0), (0.0, 10.0), (64, 64)
    # `tower_height`: an int (tower task, trapezoid library) or a task_spec() dict
    t = tower_height if isinstance(tower_height, dict) else dict(task_def(tower_height), shapes=["trapezoid"])
    env = AssemblyGym(shapes=[Shape(urdf_file=f"shapes/{nm}.urdf", name=nm) for nm in t["shapes"]], obstacles=t["obstacles"],
                      targets=t["targets"], reward_fct=sparse_reward, restrict_2d=True, max_steps=max_steps,
                      assembly_env=AssemblyEnv())
    steps, timed, warm = 0, 0.0, 0
    t_begin = time.perf_counter()
    t_end = t_begin + budget_s

    def more():
        now = time.perf_counter()
        return now < t_end and (steps < max_env_steps or now - t_begin < min_s)
    while more():
        obs, _ = env.reset()
        obstacle_f = render_blocks_2d(obs['obstacle_blocks'], xlim, ylim, img).astype(np.float32)[None]
        done = False
        while not done and more():
            block_f, _ = ofeat.get_state_features(obs, xlim, ylim, img)
            cands = [*oact.generate_actions(env, X_GROUND, [0.0])]
            cand_f = ofeat.get_action_features(env, cands, xlim, ylim, img)
            kept, _, _ = oact.filter_actions(env, cands, cand_f, block_f, obstacle_f, xlim, ylim)
            if not kept:
                break
            action = kept[int(rng.integers(len(kept)))]
            t0 = time.perf_counter()
            obs, reward, terminated, truncated, _ = env.step(action)
            env.stabilities_freezing()
            ofeat.get_state_features(obs, xlim, ylim, img)
            if warm < warm_env_steps:          # untimed warm-up (first HiGHS / numpy calls of the process)
                warm += 1
            else:
                timed += time.perf_counter() - t0
                steps += 1
            done = bool(terminated or truncated)
    return steps, timed


def cpu_env_rate(cores, budget_s, tower_height, max_steps, max_env_steps=10 ** 9, warm_env_steps=4, min_s=0.0):
    """Aggregate env steps/s of `cores` independent oracle envs (the first `warm_env_steps` env steps of every
    worker are neither counted nor timed; a worker runs until it has done `max_env_steps` AND `min_s` seconds,
    at most `budget_s`)."""
    jobs = [(1000 + i, budget_s, tower_height, max_steps, max_env_steps, warm_env_steps, min_s) for i in range(cores)]
    if cores == 1:
        results = [_cpu_worker(jobs[0])]
    else:
        with mp.get_context("spawn").Pool(cores) as pool:
            results = pool.map(_cpu_worker, jobs)
    steps = sum(r[0] for r in results)
    rate = sum(r[0] / r[1] for r in results if r[1] > 0)
    return rate, steps


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = args.cpu_cores or os.cpu_count() or 1
    # one "step" of this arm = a bounded sample: every worker advances 4 env steps; W warm-up steps per worker
    # are untimed, and the timed sample lasts at least 8 s (a shorter one measures process start-up, not the env)
    per_step = 4 * cores
    total_env_steps = per_step * args.steps
    t0 = time.perf_counter()
    rate, steps = cpu_env_rate(cores, 150.0, task_spec(args), args.max_steps,
                               max_env_steps=max(8, total_env_steps // cores),
                               warm_env_steps=min(64, 4 * max(1, args.warmup)), min_s=8.0)
    wall = time.perf_counter() - t0
    k_done = max(1, min(args.steps, steps // per_step))
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": world, "steps": k_done,
            "warmup": args.warmup, "ms_per_step": 1e3 * per_step / rate if rate > 0 else None,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args, world),
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{steps} env steps of the restated reference env (oracle/, numpy + HiGHS; "
                                       f"reference unbuildable here: compas/compas_cra/pyomo/ipopt absent) in "
                                       f"{wall:.1f} s wall on {cores} processes"},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------ GPU arm
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in out.strip().splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def profiled_traffic():
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        try:
            return json.load(open(path)).get("step_kernel_dram_bytes_per_launch")
        except Exception:
            return None
    return None


def run_sweep(args, rank, local_rank, world, dev):
    """BASELINE.json configs[3]: 65,536 random assemblies (1..15 blocks, shapes trapezoid / hexagon /
    cube1, mu cycled over 0.3 / 0.8 / 2.0), sharded contiguously over the ranks; the timed unit is
    one stability check of every assembly (interfaces + both equilibrium verdicts), no placement."""
    import numpy as np
    import torch
    from bridges_b200.envs.batched import BatchedAssemblyGym
    from bridges_b200.sharding import max_over_ranks, shard_range
    total = args.sweep_assemblies
    lo, hi = shard_range(total, rank, world)
    n = hi - lo
    env = BatchedAssemblyGym(n, ["shapes/trapezoid.urdf", "shapes/hexagon.urdf", "shapes/cube1.urdf"], max_steps=None,
                             device=local_rank)
    ids = np.arange(lo, hi)
    env.set_mu(np.array([0.3, 0.8, 2.0])[ids % 3])
    env.reset(dict())
    rng = np.random.default_rng(0)
    target = rng.integers(1, 16, size=total)[lo:hi]                 # n_blocks ~ U{1..15}
    for k in range(15):
        env.enumerate_actions(np.linspace(-2.0, 4.0, 13), (0.0, 0.25, -0.25), amax=1024, with_bits=False)
        acts, _ = env.select_random(seed=12345 + k)
        env.step(acts, mask=(target > k).astype(np.uint8))
    env.sync()
    K = args.sweep_steps
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    for _ in range(3):
        env.evaluate()
    torch.cuda.synchronize()
    for a, b in ev:
        a.record()
        env.evaluate()
        b.record()
    torch.cuda.synchronize()
    t = sum(a.elapsed_time(b) for a, b in ev) * 1e-3
    out = env.read_out()
    (t,) = max_over_ranks([t], device=dev)
    stats = dict(mean_blocks=float(out["n_blocks"].mean()), stable_frac=float(out["stable"].mean()),
                 stable_unfrozen_frac=float(out["stable_unfrozen"].mean()),
                 mean_newton_iters=float(out["newton_iters"].mean()),
                 not_converged=int(((out["solver_status"] & 3) != 0).sum()))
    env.close()
    return {"metric": "assembly stability checks/sec (two verdicts each)", "value": total * K / t,
            "unit": "assemblies/s", "assemblies": total, "steps": K, "ms_per_pass": 1e3 * t / K, "scaling": "strong",
            "rank0_stats": stats}


def run_batch_scan(args, local_rank, sizes):
    """The same step at larger lock-step batches (device-timed, L2 flushed between steps): at 1024
    envs a launch is a partial wave whose length is the slowest environment's solve; with more
    environments per GPU the SMs fill up and the per-step cost approaches the mean solve."""
    import numpy as np
    import torch
    from bridges_b200.envs.batched import BatchedAssemblyGym
    dev = torch.device("cuda", local_rank)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    rows = []
    for E in sizes:
        env = BatchedAssemblyGym(E, ["shapes/trapezoid.urdf"], max_steps=args.max_steps, device=local_rank)
        env.reset(task_def(args.tower_height))
        img = torch.empty((E, 1, 64, 64), dtype=torch.float32, device=dev)
        binary = torch.empty((E, 6), dtype=torch.float32, device=dev)
        K = 60
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        n_pre = []
        for i in range(20 + K):
            env.enumerate_actions(X_GROUND, (0.0,), amax=128, with_bits=False)
            acts = env.select_random(seed=args.seed * 1000003 + i * 7919 + 17)[0]
            if i >= 20:
                flush.fill_(i & 0xff)
                ev[i - 20][0].record()
            env.step(acts, block_img=img, binary=binary)
            if i >= 20:
                ev[i - 20][1].record()
                if i % 20 == 0:
                    n_pre.append(np.maximum(env.read_out()["n_blocks"].astype(np.float64) - 1, 0).mean())
            env.reset_done()
        torch.cuda.synchronize()
        t = sum(a.elapsed_time(b) for a, b in ev) * 1e-3
        bytes_per_launch = E * (16528 + 16 * args.tower_height + 32 * float(np.mean(n_pre)))
        rows.append({"envs": E, "env_steps_per_s": E * K / t, "ms_per_step": 1e3 * t / K,
                     "achieved_GBps": bytes_per_launch * K / t / 1e9})
        env.close()
        del env, img, binary
    return rows


def run_gpu(args, rank, local_rank, world):
    import numpy as np
    import torch
    import torch.distributed as dist
    from bridges_b200 import lib as L
    from bridges_b200.envs.batched import BatchedAssemblyGym
    from bridges_b200.sharding import max_over_ranks

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: bridges_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_cpus = None
    if world > 1 and not args.no_numa_bind:
        # several ranks share the host: keep this rank (and the pinned buffers it allocates) on its GPU's NUMA node
        from bridges_b200.sharding import bind_to_gpu_numa_node
        numa_cpus = bind_to_gpu_numa_node(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # CPU baseline first (rank 0, N=1 only), before this process touches CUDA in earnest
    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = args.cpu_cores or os.cpu_count() or 1
        t0 = time.perf_counter()
        rate, steps = cpu_env_rate(cores, args.cpu_budget, task_spec(args), args.max_steps)
        rate1, steps1 = cpu_env_rate(1, min(5.0, args.cpu_budget), task_spec(args), args.max_steps)
        cpu_base = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                    "one_core": {"value": rate1, "unit": UNIT, "sample": f"{steps1} env steps in one process"},
                    "sample": f"{steps} env steps of the restated reference CPU env (oracle/: numpy + HiGHS, "
                              f"reference call pattern 5 solves + 3 interface rebuilds + 1 raster per step) in "
                              f"{time.perf_counter() - t0:.1f} s wall on {cores} processes; same task and policy"}

    E, K, W = args.envs, args.steps, args.warmup
    spec = task_spec(args)
    env = BatchedAssemblyGym(E, [f"shapes/{nm}.urdf" for nm in spec["shapes"]], max_steps=args.max_steps, device=local_rank)
    lib, h = env.lib, env.handle
    env.reset(dict(obstacles=spec["obstacles"], targets=spec["targets"]))
    n_obs_tgt = len(spec["obstacles"]) + len(spec["targets"])
    dt = env.dt
    block_img = torch.empty((E, 1, 64, 64), dtype=torch.float32, device=dev)
    binary = torch.empty((E, 6), dtype=torch.float32, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    # capacity of the candidate list: (sum of faces over the library) x (10 ground offsets + free receiver faces)
    # (trapezoid: 44 + 8 n candidates with n placed blocks)
    amax = (128 if args.max_steps <= 10 else 256) if spec["shapes"] == ["trapezoid"] else 1024
    lib.bw_set_timing(h, 0)

    cand_ev = []

    def choose_actions(step_id, timed=False):
        if timed:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
        env.enumerate_actions(X_GROUND, (0.0,), amax=amax, with_bits=True)
        acts = env.select_random(seed=args.seed * 1000003 + step_id * 7919 + rank)[0]
        if timed:
            b.record()
            cand_ev.append((a, b))
        return acts

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (also brings the env population to its steady-state mix of block counts)
    for i in range(W):
        acts = choose_actions(i)
        env.step(acts, block_img=block_img, binary=binary)
        env.reset_done()
    barrier()

    # ---- device-timed loop
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    stats = []
    sampler = ClockSampler(local_rank)
    launches0 = env.kernel_launches()
    sampler.start()
    barrier()
    wall0 = time.perf_counter()
    for i in range(K):
        acts = choose_actions(W + i, timed=True)
        if not args.no_flush:
            flush.fill_(i & 0xff)                   # evict the 126 MB L2
        ev[i][0].record()
        env.step(acts, block_img=block_img, binary=binary)
        ev[i][1].record()
        if i % max(1, K // 16) == 0:
            stats.append(env._out.clone())
        env.reset_done()
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop()
    launches = env.kernel_launches() - launches0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    t_dev = sum(step_ms) * 1e-3
    t_cand = sum(a.elapsed_time(b) for a, b in cand_ev) * 1e-3

    # ---- statistics of the sampled steps (algorithmic bytes / flops per launch)
    outs = np.concatenate([s.cpu().numpy().view(dt["step_out"]) for s in stats])
    n_pre = np.maximum(outs["n_blocks"].astype(np.float64) - 1, 0)
    bytes_per_launch = float(E * (16528 + 16 * n_obs_tgt) + 32 * n_pre.mean() * E)
    flops_per_launch = float(outs["solver_kflops"].astype(np.float64).mean() * 1e3 * E)
    fp64 = C.c_double(0.0)
    lib.bw_fp64_peak_gflops(h, C.byref(fp64))

    # ---- end-to-end loops through the host entry point bw_step_host: pinned host actions in,
    # host results out.  (a) headline: step records + the raster as the reference's `_get_obs` /
    # `render_blocks_2d` deliver it (one byte per pixel, bool) + binary features;
    # (b) the same with the float32 [E,1,64,64] tensor of get_state_features copied to the host too.
    h_act = torch.empty(E * dt["action"].itemsize, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(E * dt["step_out"].itemsize, dtype=torch.uint8).pin_memory()
    h_u8 = torch.empty((E, 64, 64), dtype=torch.uint8).pin_memory()
    h_img = torch.empty((E, 1, 64, 64), dtype=torch.float32).pin_memory()
    h_bin = torch.empty((E, 6), dtype=torch.float32).pin_memory()
    h_bits = torch.empty((E, 64), dtype=torch.int64).pin_memory()
    Ke = min(K, args.e2e_steps)

    def e2e_loop(obs, seed0):
        barrier()
        total = 0.0
        for i in range(Ke):
            acts = choose_actions(seed0 + i)
            h_act.copy_(acts)                       # the policy's actions arrive on the host
            if not args.no_flush:
                flush.fill_(i & 0xff)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            rc = lib.bw_step_host(h, h_act.data_ptr(), None, h_out.data_ptr(), C.byref(obs))
            total += time.perf_counter() - t0
            L.check(lib, h, rc)
            env.reset_done()
        barrier()
        return total

    t_e2e = e2e_loop(L.bw_obs_out(None, h_u8.data_ptr(), h_bin.data_ptr()), W + K)
    t_e2e_f32 = e2e_loop(L.bw_obs_out(h_img.data_ptr(), None, h_bin.data_ptr()), W + K + Ke)
    # the same call with the zero-copy path switched off: H2D memcpy, kernel, D2H memcpys, in sequence
    lib.bw_set_host_transfer(h, 1)
    t_e2e_staged = e2e_loop(L.bw_obs_out(None, h_u8.data_ptr(), h_bin.data_ptr()), W + K + 2 * Ke)
    lib.bw_set_host_transfer(h, 0)
    # (c) the raster bit-packed (bw_obs_out.block_bits, 512 B per environment) for host consumers that unpack lazily
    t_e2e_bits = e2e_loop(L.bw_obs_out(None, None, h_bin.data_ptr(), h_bits.data_ptr()), W + K + 3 * Ke)
    d2h_bits = E * (dt["step_out"].itemsize + 64 * 8 + 6 * 4)
    h2d = E * dt["action"].itemsize
    d2h = E * (dt["step_out"].itemsize + 64 * 64 + 6 * 4)
    d2h_f32 = E * (dt["step_out"].itemsize + 64 * 64 * 4 + 6 * 4)

    # ---- max over ranks
    sweep = run_sweep(args, rank, local_rank, world, dev) if args.sweep else None
    t_dev, t_e2e, t_e2e_f32, wall, t_cand, t_e2e_staged, t_e2e_bits = max_over_ranks(
        [t_dev, t_e2e, t_e2e_f32, wall, t_cand, t_e2e_staged, t_e2e_bits], device=dev)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peaks()
    ms_kernel = 1e3 * t_dev / K
    achieved = bytes_per_launch / (ms_kernel * 1e-3) / 1e9
    value = world * E * K / t_dev
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_kernel, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": config_dict(args, world),
        "clocks": clocks,
        "e2e": {"value": world * E * Ke / t_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": Ke, "api": "bw_step_host: pinned host actions in; step records + u8 raster [E,64,64] "
                                    "(render_blocks_2d's bool image) + binary features out; pinned buffers are "
                                    "read / written by the kernel over PCIe as each environment finishes",
                "staged_copies": {"value": world * E * Ke / t_e2e_staged, "unit": UNIT,
                                  "note": "same call, bw_set_host_transfer(h, 1): cudaMemcpyAsync before and after the kernel"},
                "with_f32_images": {"value": world * E * Ke / t_e2e_f32, "unit": UNIT, "d2h_bytes_per_step": d2h_f32},
                "with_bit_rasters": {"value": world * E * Ke / t_e2e_bits, "unit": UNIT, "d2h_bytes_per_step": d2h_bits,
                                     "note": "not the headline: the raster leaves the GPU bit-packed (64 x u64 per "
                                             "environment), BatchedAssemblyGym.bits_to_bool unpacks it on demand"}},
        "gpu_launches": int(launches),
        "roofline": {"kernel": "step_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": profiled_traffic(), "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": bytes_per_launch,
                     "note": "latency/FP64-bound kernel: see roofline_fp64 for the solver"},
        "roofline_fp64": {"kernel": "step_kernel (equilibrium solves)", "bound": "fp64-fma",
                          "achieved": flops_per_launch / (ms_kernel * 1e-3) / 1e12, "peak": fp64.value / 1e3,
                          "unit": "TFLOP/s", "frac": (flops_per_launch / (ms_kernel * 1e-3) / 1e9) / max(fp64.value, 1e-9),
                          "peak_source": "bw_fp64_peak_gflops micro-benchmark, same job",
                          "flops_per_launch": flops_per_launch},
        "env_stats": {"mean_blocks": float(outs["n_blocks"].mean()), "mean_interfaces": float(outs["n_interfaces"].mean()),
                      "mean_newton_iters_per_step": float(outs["newton_iters"].mean()),
                      "stable_frac": float(outs["stable"].mean()), "terminated_frac": float(outs["terminated"].mean()),
                      "solver_not_converged": int(((outs["solver_status"] & 3) != 0).sum()),
                      "verdicts_implied_frac": float(((outs["solver_status"] & 4) != 0).mean() + ((outs["solver_status"] & 8) != 0).mean()) / 2},
        "with_candidate_stage": {"value": world * E * K / (t_dev + t_cand), "unit": UNIT,
                                 "candidate_ms_per_step": 1e3 * t_cand / K,
                                 "note": "step + enumerate/filter kernel (candidates with bit rasters and validity "
                                         "mask, robotoddler/utils/actions.py:7-82) + random selection"},
        "wall_s_timed_region": wall,
    }
    line["e2e"]["host_binding"] = (f"rank bound to the {len(numa_cpus)} CPUs next to its GPU" if numa_cpus else "none")
    if sweep is not None:
        line["sweep"] = sweep
    if args.batch_scan:
        scan = run_batch_scan(args, local_rank, [4096, 16384, 65536])
        peak_hbm = peak
        for r in scan:
            r["hbm_frac"] = r["achieved_GBps"] / peak_hbm
        line["batch_scan"] = {"note": "same step kernel and workload at larger lock-step batches per GPU (rank 0)",
                              "rows": scan}
    if cpu_base is not None:
        line["cpu_baseline"] = cpu_base
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--envs", type=int, default=1024, help="lock-step envs per GPU")
    ap.add_argument("--tower-height", type=int, default=2)
    ap.add_argument("--task", default="tower", choices=["tower", "bridge"],
                    help="tower: the tower_height task (headline); bridge: horizontal_bridge_setup (configs[4] shape)")
    ap.add_argument("--num-obstacles", type=int, default=5, help="--task bridge: obstacle cubes to span")
    ap.add_argument("--shapes", default="trapezoid,hexagon", help="--task bridge: block library (names under shapes/)")
    ap.add_argument("--max-steps", type=int, default=10)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--cpu-budget", type=float, default=15.0, help="seconds of CPU-baseline sampling")
    ap.add_argument("--e2e-steps", type=int, default=200)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-cores", type=int, default=0, help="CPU-arm worker processes (0 = all host cores)")
    ap.add_argument("--no-numa-bind", action="store_true", help="N>1: do not bind the rank to the CPUs of its GPU's NUMA node")
    ap.add_argument("--no-flush", action="store_true", help="profiling only: skip the L2 flush between steps")
    ap.add_argument("--sweep", action="store_true", help="also run the 65,536-assembly stability sweep (configs[3])")
    ap.add_argument("--batch-scan", action="store_true", help="also time the step at 4096 / 16384 / 65536 envs per GPU")
    ap.add_argument("--sweep-assemblies", type=int, default=65536)
    ap.add_argument("--sweep-steps", type=int, default=10)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_gpu(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
