#!/usr/bin/env python
"""bench.py -- stability-checked env steps/sec of the batched assembly_gym step on B200.

    python bench.py --gpus N --steps K --warmup W                    # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W   # restated reference CPU env

Default job (what the driver runs), one JSON line on rank 0:
  * headline (`value`, `e2e`, `roofline`, ...): the heaviest env-loop configuration of BASELINE.json --
    `horizontal_bridge_setup(num_obstacles=5)` with the mixed trapezoid + hexagon library, max_steps = 15 (configs[4]
    shape), 1024 lock-step envs per GPU, uniformly random valid actions from the candidate kernel, auto-reset;
  * `secondary`: the same measurement on configs[1] (tower_height=2, max_steps=10) and configs[2] (tower_height=4,
    max_steps=15);
  * `sweep`: configs[3], 65,536 random assemblies of up to 15 blocks, sharded over the ranks (strong scaling);
  * `steady_state` (in every record): >= 1 s of timed steps with mean / median / p99 per step -- a launch lasts as long
    as its slowest environment, so a 20-step sample is noisy;
  * `rollout`: the fused rollout (`bw_rollout_random`) with its packed transition records; at N > 1 every chunk is
    all-gathered into the replay ring of every rank over NCCL on a side stream (`replay_gather`);
  * `parity_gate`: a short lock-step run of the timed path replayed through the CPU oracle in the same job.

One "step" = one `bw_step` over the whole batch: block placement, contact interfaces, the two equilibrium verdicts
(new block frozen / released), targets, reward, termination, raster update and the f32 [E,1,64,64] observation +
[E,6] binary features.  Timing: every step is bracketed by CUDA events on the stream the
kernel is launched on; the L2 is flushed (256 MiB write) between steps, outside the event pairs; the loop is
bracketed by barrier + synchronize and the max over ranks is reported.
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
CAND_BITS = True if os.environ.get("BW_BENCH_CAND_BITS", "stored") == "dense" else "stored"   # rasters of the candidate stage

# the env-loop workloads of BASELINE.json
WORKLOADS = {
    "bridge": dict(task="bridge", num_obstacles=5, shapes="trapezoid,hexagon", tower_height=2, max_steps=15),
    "tower2": dict(task="tower", num_obstacles=5, shapes="trapezoid", tower_height=2, max_steps=10),
    "tower4": dict(task="tower", num_obstacles=5, shapes="trapezoid", tower_height=4, max_steps=15),
}
HEADLINE, SECONDARY = "bridge", ("tower2", "tower4")


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
    common = {"envs_per_gpu": args.envs, "max_steps": args.max_steps, "image": "64x64 f32", "mu": 0.8,
              "parallelism": f"env-sharded x{n_gpus}, no collective on the step path",
              "cache": "L2 flushed between timed steps (256 MiB write outside the event pairs)"}
    if args.task == "bridge":
        return dict(common, workload=f"horizontal_bridge_setup(num_obstacles={args.num_obstacles}) with shapes "
                                     f"[{args.shapes}], {args.envs} lock-step envs per GPU, max_steps={args.max_steps}, "
                                     f"random valid actions, auto-reset (BASELINE.json configs[4] shape)",
                    num_obstacles=args.num_obstacles, shapes=args.shapes)
    tag = "configs[1]" if (args.tower_height, args.max_steps) == (2, 10) else "configs[2] shape"
    return dict(common, workload=f"tower_height={args.tower_height} trapezoid task, {args.envs} lock-step envs per GPU, "
                                 f"max_steps={args.max_steps}, random valid actions, auto-reset (BASELINE.json {tag})",
                tower_height=args.tower_height)


def workload_args(args, name):
    """argparse.Namespace of one named workload (the other options are inherited)."""
    ns = argparse.Namespace(**vars(args))
    for k, v in WORKLOADS[name].items():
        setattr(ns, k, v)
    return ns


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


def cpu_baseline_record(args, budget_s, with_one_core=True):
    cores = args.cpu_cores or os.cpu_count() or 1
    t0 = time.perf_counter()
    rate, steps = cpu_env_rate(cores, budget_s, task_spec(args), args.max_steps)
    rec = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
           "sample": f"{steps} env steps of the restated reference CPU env (oracle/: numpy + HiGHS, reference call "
                     f"pattern 5 solves + 3 interface rebuilds + 1 raster per step) in {time.perf_counter() - t0:.1f} s "
                     f"wall on {cores} processes; same task and policy"}
    if with_one_core:
        rate1, steps1 = cpu_env_rate(1, min(5.0, budget_s), task_spec(args), args.max_steps)
        rec["one_core"] = {"value": rate1, "unit": UNIT, "sample": f"{steps1} env steps in one process"}
    return rec


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


def profiled_traffic(key="step_kernel_dram_bytes_per_launch"):
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        try:
            return json.load(open(path)).get(key)
        except Exception:
            return None
    return None


class Job:
    """What every measurement of one rank shares: device, flush buffer, barrier, reductions."""

    def __init__(self, args, rank, local_rank, world):
        import torch
        import torch.distributed as dist
        self.args, self.rank, self.local_rank, self.world = args, rank, local_rank, world
        self.torch, self.dist = torch, dist
        self.dev = torch.device("cuda", local_rank)
        self.flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=self.dev)
        self.peak, self.peak_src = measured_peaks()
        self.fp64_gflops = None

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, values):
        from bridges_b200.sharding import max_over_ranks
        return max_over_ranks(values, device=self.dev)


def measure_workload(job, wargs, K, W, steady_seconds, with_e2e_variants=True):
    """Device-timed step, steady-state sample, candidate stage, end-to-end host path and the two rooflines of one
    env-loop workload.  Collective (every rank calls it); returns the record on every rank."""
    import numpy as np
    torch = job.torch
    from bridges_b200 import lib as L
    from bridges_b200.envs.batched import BatchedAssemblyGym
    args, dev, world, rank = job.args, job.dev, job.world, job.rank
    E = wargs.envs
    spec = task_spec(wargs)
    env = BatchedAssemblyGym(E, [f"shapes/{nm}.urdf" for nm in spec["shapes"]], max_steps=wargs.max_steps, device=job.local_rank)
    lib, h = env.lib, env.handle
    env.reset(dict(obstacles=spec["obstacles"], targets=spec["targets"]))
    n_obs_tgt = len(spec["obstacles"]) + len(spec["targets"])
    dt = env.dt
    block_img = torch.empty((E, 1, 64, 64), dtype=torch.float32, device=dev)
    binary = torch.empty((E, 6), dtype=torch.float32, device=dev)
    flush = job.flush
    # capacity of the candidate list: (sum of target faces over the library) x (10 ground offsets + receiver faces)
    amax = min(1024, max(128, 64 * ((env.max_candidates(len(X_GROUND), 1) + 63) // 64)))
    lib.bw_set_timing(h, 0)
    cand_ev = []

    def choose_actions(step_id, timed=False):
        if timed:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
        env.enumerate_actions(X_GROUND, (0.0,), amax=amax, with_bits=CAND_BITS)
        acts = env.select_random(seed=args.seed * 1000003 + step_id * 7919 + rank)[0]
        if timed:
            b.record()
            cand_ev.append((a, b))
        return acts

    # ---- warm-up (also brings the env population to its steady-state mix of block counts).  Everything that makes
    # the GPU wait -- the first NCCL collective (communicator set-up), the start of the nvidia-smi sampler -- happens
    # BEFORE it, so that the barrier in front of the timed steps is a few microseconds and the clocks stay up
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    stats = []
    sampler = ClockSampler(job.local_rank)
    job.barrier()
    sampler.start()
    for i in range(max(W, 3 * wargs.max_steps)):
        acts = choose_actions(i)
        env.step(acts, block_img=block_img, binary=binary)
        env.reset_done()
    launches0 = env.kernel_launches()

    # ---- device-timed loop: exactly K steps
    job.barrier()
    wall0 = time.perf_counter()
    for i in range(K):
        acts = choose_actions(W + i, timed=True)
        if not args.no_flush:
            flush.fill_(i & 0xff)                   # evict the 126 MB L2
        ev[i][0].record()
        env.step(acts, block_img=block_img, binary=binary)
        ev[i][1].record()
        if i % max(4, K // 16) == 0:
            stats.append(env._out.clone())
        env.reset_done()
    job.barrier()
    wall = time.perf_counter() - wall0
    launches = env.kernel_launches() - launches0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    t_dev = sum(step_ms) * 1e-3
    t_cand = sum(a.elapsed_time(b) for a, b in cand_ev) * 1e-3

    # ---- steady state: timed steps until >= steady_seconds of device time (bounded), per-step distribution
    ss_ms = []
    if steady_seconds > 0:
        chunk, total, sid = 256, 0.0, W + K
        while len(ss_ms) < 40000:
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(chunk)]
            for i in range(chunk):
                acts = choose_actions(sid + i)
                if not args.no_flush:
                    flush.fill_(i & 0xff)
                evs[i][0].record()
                env.step(acts, block_img=block_img, binary=binary)
                evs[i][1].record()
                if i % 64 == 0:
                    stats.append(env._out.clone())
                env.reset_done()
            torch.cuda.synchronize()
            ms = [a.elapsed_time(b) for a, b in evs]
            ss_ms.extend(ms)
            total += sum(ms) * 1e-3
            sid += chunk
            # every rank runs the same number of chunks: continue while ANY rank is short of its sample
            (short,) = job.max_over_ranks([1.0 if total < steady_seconds else 0.0])
            if short == 0.0:
                break
        job.barrier()
    clocks = sampler.stop()

    # ---- statistics of the sampled steps (algorithmic bytes / flops per launch)
    outs = np.concatenate([s.cpu().numpy().view(dt["step_out"]) for s in stats])
    n_pre = np.maximum(outs["n_blocks"].astype(np.float64) - 1, 0)
    bytes_per_launch = float(E * (16528 + 16 * n_obs_tgt) + 32 * n_pre.mean() * E)
    flops_per_launch = float(outs["solver_kflops"].astype(np.float64).mean() * 1e3 * E)
    if job.fp64_gflops is None:
        fp64 = C.c_double(0.0)
        lib.bw_fp64_peak_gflops(h, C.byref(fp64))
        job.fp64_gflops = fp64.value

    # ---- end-to-end loops through the host entry point bw_step_host: pinned host actions in, host results out.
    # headline: step records + the raster bit-packed (64 x u64 per environment: `_get_obs` returns blocks, not images;
    # BatchedAssemblyGym.bits_to_bool expands on demand) + binary features; secondary: the raster as one byte per
    # pixel (render_blocks_2d's bool image), as the float32 tensor of get_state_features, and with staged copies.
    h_act = torch.empty(E * dt["action"].itemsize, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(E * dt["step_out"].itemsize, dtype=torch.uint8).pin_memory()
    h_bin = torch.empty((E, 6), dtype=torch.float32).pin_memory()
    h_bits = torch.empty((E, 64), dtype=torch.int64).pin_memory()
    Ke = max(1, min(max(K, 50), args.e2e_steps))

    def e2e_loop(obs, seed0):
        job.barrier()
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
        job.barrier()
        return total

    base_seed = W + K + 100000
    t_e2e = e2e_loop(L.bw_obs_out(None, None, h_bin.data_ptr(), h_bits.data_ptr()), base_seed)
    h2d = E * dt["action"].itemsize
    d2h = E * (dt["step_out"].itemsize + 64 * 8 + 6 * 4)
    variants = {}
    if with_e2e_variants:
        h_u8 = torch.empty((E, 64, 64), dtype=torch.uint8).pin_memory()
        h_img = torch.empty((E, 1, 64, 64), dtype=torch.float32).pin_memory()
        t_u8 = e2e_loop(L.bw_obs_out(None, h_u8.data_ptr(), h_bin.data_ptr()), base_seed + Ke)
        t_f32 = e2e_loop(L.bw_obs_out(h_img.data_ptr(), None, h_bin.data_ptr()), base_seed + 2 * Ke)
        lib.bw_set_host_transfer(h, 1)              # zero-copy off: H2D memcpy, kernel, D2H memcpys, in sequence
        t_staged = e2e_loop(L.bw_obs_out(None, None, h_bin.data_ptr(), h_bits.data_ptr()), base_seed + 3 * Ke)
        lib.bw_set_host_transfer(h, 0)
        t_u8, t_f32, t_staged = job.max_over_ranks([t_u8, t_f32, t_staged])
        variants = {
            "with_u8_images": {"value": world * E * Ke / t_u8, "unit": UNIT,
                               "d2h_bytes_per_step": E * (dt["step_out"].itemsize + 64 * 64 + 6 * 4),
                               "note": "the raster as render_blocks_2d's bool image, one byte per pixel [E,64,64]"},
            "with_f32_images": {"value": world * E * Ke / t_f32, "unit": UNIT,
                                "d2h_bytes_per_step": E * (dt["step_out"].itemsize + 64 * 64 * 4 + 6 * 4)},
            "staged_copies": {"value": world * E * Ke / t_staged, "unit": UNIT,
                              "note": "same call as the headline, bw_set_host_transfer(h, 1): cudaMemcpyAsync before "
                                      "and after the kernel"}}

    t_dev, t_e2e, wall, t_cand = job.max_over_ranks([t_dev, t_e2e, wall, t_cand])
    ss = None
    if ss_ms:
        srt = sorted(ss_ms)
        mean_ms, med_ms, p99_ms = sum(ss_ms) / len(ss_ms), srt[len(srt) // 2], srt[min(len(srt) - 1, int(0.99 * len(srt)))]
        mean_ms, med_ms, p99_ms = job.max_over_ranks([mean_ms, med_ms, p99_ms])
        ss = {"steps": len(ss_ms), "seconds": sum(ss_ms) * 1e-3, "mean_ms": mean_ms, "median_ms": med_ms, "p99_ms": p99_ms,
              "value": world * E / (mean_ms * 1e-3), "unit": UNIT,
              "note": "per-step CUDA-event times of a >= %.1f s sample (max over ranks of each statistic)" % steady_seconds}
    env.close()

    ms_kernel = 1e3 * t_dev / K
    achieved = bytes_per_launch / (ms_kernel * 1e-3) / 1e9
    rec = {
        "value": world * E * K / t_dev, "unit": UNIT, "steps": K, "ms_per_step": ms_kernel,
        "config": config_dict(wargs, world), "clocks": clocks,
        "e2e": dict({"value": world * E * Ke / t_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                     "steps": Ke,
                     "api": "bw_step_host: pinned host actions in; step records + bit-packed raster [E,64] u64 "
                            "(the adapter expands it lazily: `_get_obs` returns blocks, gym_env.py:179-195) + binary "
                            "features out; pinned buffers are read / written by the kernel over PCIe as each "
                            "environment finishes"}, **variants),
        "gpu_launches": int(launches),
        "roofline": {"kernel": "step_kernel", "bound": "hbm", "achieved": achieved, "peak": job.peak, "unit": "GB/s",
                     "frac": achieved / job.peak, "traffic": profiled_traffic("step_kernel_dram_bytes_per_launch_" + wargs.task)
                     or profiled_traffic(), "peak_source": job.peak_src, "algorithmic_bytes_per_launch": bytes_per_launch,
                     "note": "latency / FP64-bound kernel (a launch lasts as long as its slowest environment's solve): "
                             "see roofline_fp64 and steady_state"},
        "roofline_fp64": {"kernel": "step_kernel (equilibrium solves)", "bound": "fp64-fma",
                          "achieved": flops_per_launch / (ms_kernel * 1e-3) / 1e12, "peak": job.fp64_gflops / 1e3,
                          "unit": "TFLOP/s", "frac": (flops_per_launch / (ms_kernel * 1e-3) / 1e9) / max(job.fp64_gflops, 1e-9),
                          "peak_source": "bw_fp64_peak_gflops micro-benchmark, same job",
                          "flops_per_launch": flops_per_launch},
        "env_stats": {"mean_blocks": float(outs["n_blocks"].mean()), "mean_interfaces": float(outs["n_interfaces"].mean()),
                      "mean_newton_iters_per_step": float(outs["newton_iters"].mean()),
                      "stable_frac": float(outs["stable"].mean()), "terminated_frac": float(outs["terminated"].mean()),
                      "solver_not_converged": int(((outs["solver_status"] & 3) != 0).sum()),
                      "verdicts_implied_frac": float(((outs["solver_status"] & 4) != 0).mean() + ((outs["solver_status"] & 8) != 0).mean()) / 2,
                      "verdicts_by_lp_frac": float(((outs["solver_status"] & 16) != 0).mean() + ((outs["solver_status"] & 32) != 0).mean()) / 2,
                      "mean_lp_pivots_per_step": float(outs["lp_pivots"].mean()), "max_lp_pivots": int(outs["lp_pivots"].max())},
        "with_candidate_stage": {"value": world * E * K / (t_dev + t_cand), "unit": UNIT,
                                 "candidate_ms_per_step": 1e3 * t_cand / K, "amax": amax,
                                 "candidate_rasters": "store slots" if CAND_BITS == "stored" else "dense copies",
                                 "note": "step + enumerate/filter kernel (candidates, validity mask and their bit "
                                         "rasters -- kept in the handle's candidate store and handed out as slots, "
                                         "BW_BENCH_CAND_BITS=dense: copied out as [E,amax,64] -- "
                                         "robotoddler/utils/actions.py:7-82) + random selection: the rate a rollout sees"},
        "wall_s_timed_region": wall,
    }
    if ss is not None:
        rec["steady_state"] = ss
    return rec


def run_sweep(job):
    """BASELINE.json configs[3]: 65,536 random assemblies (1..15 blocks, shapes trapezoid / hexagon / cube1, mu
    cycled over 0.3 / 0.8 / 2.0), sharded contiguously over the ranks; the timed unit is one stability check of
    every assembly (interfaces + both equilibrium verdicts), no placement."""
    import numpy as np
    torch = job.torch
    from bridges_b200.envs.batched import BatchedAssemblyGym
    from bridges_b200.sharding import shard_range
    args, rank, world, dev = job.args, job.rank, job.world, job.dev
    total = args.sweep_assemblies
    lo, hi = shard_range(total, rank, world)
    n = hi - lo
    env = BatchedAssemblyGym(n, ["shapes/trapezoid.urdf", "shapes/hexagon.urdf", "shapes/cube1.urdf"], max_steps=None,
                             device=job.local_rank)
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
    job.barrier()
    for a, b in ev:
        a.record()
        env.evaluate()
        b.record()
    torch.cuda.synchronize()
    t = sum(a.elapsed_time(b) for a, b in ev) * 1e-3
    out = env.read_out()
    (t,) = job.max_over_ranks([t])
    # algorithmic bytes of one check: 32 + 32 n_blocks in, 16 + 32 n_itf out
    bytes_local = float((32 + 32 * out["n_blocks"].astype(np.float64)).sum() + (16 + 32 * out["n_interfaces"].astype(np.float64)).sum())
    flops_local = float(out["solver_kflops"].astype(np.float64).sum() * 1e3)
    tot = torch.tensor([bytes_local, flops_local], dtype=torch.float64, device=dev)
    if world > 1:
        job.dist.all_reduce(tot)
    stats = dict(mean_blocks=float(out["n_blocks"].mean()), stable_frac=float(out["stable"].mean()),
                 stable_unfrozen_frac=float(out["stable_unfrozen"].mean()),
                 mean_newton_iters=float(out["newton_iters"].mean()),
                 not_converged=int(((out["solver_status"] & 3) != 0).sum()))
    env.close()
    gbs = float(tot[0]) * K / t / 1e9
    tfl = float(tot[1]) * K / t / 1e12
    return {"metric": "assembly stability checks/sec (two verdicts each)", "value": total * K / t,
            "unit": "assemblies/s", "assemblies": total, "steps": K, "ms_per_pass": 1e3 * t / K, "scaling": "strong",
            "config": {"workload": "65,536 seeded random assemblies, 1-15 blocks, shapes trapezoid/hexagon/cube1, mu cycled over "
                                   "0.3/0.8/2.0 (BASELINE.json configs[3]), sharded contiguously over the ranks"},
            "roofline": {"kernel": "step_kernel<16 blocks> (evaluation only)", "bound": "hbm", "achieved": gbs,
                         "peak": job.peak * world, "unit": "GB/s", "frac": gbs / (job.peak * world),
                         "traffic": profiled_traffic("sweep_dram_bytes_per_launch"), "peak_source": job.peak_src,
                         "note": "a stability check moves < 2 KB per assembly: the pass is bound by the solver's "
                                 "dependent FP64 chains, not by bytes"},
            "roofline_fp64": {"bound": "fp64-fma", "achieved": tfl, "peak": world * (job.fp64_gflops or 0.0) / 1e3,
                              "unit": "TFLOP/s", "frac": tfl / max(world * (job.fp64_gflops or 0.0) / 1e3, 1e-12)},
            "rank0_stats": stats}


def run_rollout(job, wargs):
    """The fused rollout on the headline task: `bw_rollout_random` in chunks of T iterations, packed 1,608-byte
    records, every chunk copied (N = 1) or all-gathered over NCCL (N > 1) into the replay ring of every rank on a
    side stream while the next chunk runs."""
    import numpy as np
    torch = job.torch
    from bridges_b200.envs.batched import BatchedAssemblyGym
    from bridges_b200.rollout import REC, FusedRollout, TransitionRing
    dev, world = job.dev, job.world
    E, T, chunks = wargs.envs, 16, 8
    spec = task_spec(wargs)
    env = BatchedAssemblyGym(E, [f"shapes/{nm}.urdf" for nm in spec["shapes"]], max_steps=wargs.max_steps, device=job.local_rank)
    env.reset(dict(obstacles=spec["obstacles"], targets=spec["targets"]))
    amax = min(1024, max(128, 64 * ((env.max_candidates(len(X_GROUND), 1) + 63) // 64)))
    ring = TransitionRing(4 * world * T * E, dev)
    roll = FusedRollout(env, X_GROUND, (0.0,), amax=amax, chunk_steps=T, ring=ring)

    def timed(gather):
        roll.collect_random(2, seed=100, gather=gather)
        roll.drain()
        job.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        roll.collect_random(chunks, seed=7, gather=gather)
        roll.drain()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) * 1e-3
    t_gather = timed(True)
    overflow = env.candidate_overflow()
    valid = float((ring.column("valid") != 0).float().mean())
    rec = ring.numpy()
    v = rec[rec["valid"] == 1]
    consistent = bool(np.array_equal(v["next_block_bits"], v["block_bits"] | v["action_bits"]))
    envs_seen = int(np.unique(rec["env"]).size)
    # the gather alone: one chunk, timed on its own
    gather_us = None
    if world > 1:
        ring_local = TransitionRing(2 * T * E, dev)               # same rollout without the collective
        roll.ring = ring_local
        t_local = timed(False)
        roll.ring = ring
        buf = roll.chunks[0]
        _, region = ring.reserve(world * buf.shape[0])
        job.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(4):
            job.dist.all_gather_into_tensor(region, buf)
        b.record()
        torch.cuda.synchronize()
        gather_us = 1e3 * a.elapsed_time(b) / 4
        t_gather, t_local, gather_us = job.max_over_ranks([t_gather, t_local, gather_us])
    else:
        (t_gather,) = job.max_over_ranks([t_gather])
    env.close()
    iters = chunks * T
    out = {"metric": "transitions/sec (fused rollout: pick -> step -> record -> auto-reset -> next candidates)",
           "value": world * E * iters / t_gather, "unit": "transitions/s", "per_gpu": E * iters / t_gather,
           "ms_per_iteration": 1e3 * t_gather / iters, "iterations": iters, "chunk_steps": T, "envs_per_gpu": E,
           "record_bytes": REC, "valid_frac": valid, "rasters_consistent": consistent, "envs_in_ring": envs_seen,
           "candidate_overflow": overflow, "config": config_dict(wargs, world)["workload"]}
    if world > 1:
        out["replay_gather"] = {"collective": "all_gather_into_tensor (NCCL) per chunk on a side stream, overlapped with "
                                              "the next chunk",
                                "bytes_received_per_chunk_per_rank": (world - 1) * T * E * REC,
                                "bytes_in_ring_per_chunk": world * T * E * REC, "us_per_chunk_alone": gather_us,
                                "busbw_GBps": (world - 1) * T * E * REC / (gather_us * 1e-6) / 1e9,
                                "transitions_per_s_without_gather": world * E * iters / t_local,
                                "gather_cost_frac": max(0.0, 1.0 - t_local / t_gather)}
    return out


def run_parity_gate(job, names, envs=32, steps=12):
    """A short run of the timed path (enumerate -> select_random -> step -> reset_done) on rank 0, replayed
    environment by environment through the CPU oracle in worker processes (tests/helpers.replay_worker): rasters,
    interface counts, verdicts outside the residual band, rewards and termination.  The oracle is the checker."""
    if job.rank != 0 or not names:
        return None
    import numpy as np
    from bridges_b200.envs.batched import BatchedAssemblyGym
    from tests import helpers as H
    band = (1e-9, 1e-4)
    out = {"envs": envs, "steps": steps, "band": list(band), "workloads": {}}
    t0 = time.perf_counter()
    for name in names:
        wargs = workload_args(job.args, name)
        spec = task_spec(wargs)
        env = BatchedAssemblyGym(envs, [f"shapes/{nm}.urdf" for nm in spec["shapes"]], max_steps=wargs.max_steps,
                                 device=job.local_rank)
        env.reset(dict(obstacles=spec["obstacles"], targets=spec["targets"]))
        amax = min(1024, max(128, 64 * ((env.max_candidates(len(X_GROUND), 1) + 63) // 64)))
        acts_log, out_log, bits_log = [], [], []
        for k in range(steps):
            env.enumerate_actions(X_GROUND, (0.0,), amax=amax, with_bits=True)
            acts, _ = env.select_random(seed=4242 + 17 * k)
            acts_log.append(acts.cpu().numpy().view(env.dt["action"]).copy())
            env.step(acts)
            out_log.append(env.read_out().copy())
            bits_log.append(env.raster_bits()[0].copy())
            env.reset_done()
        env.close()
        jobs = []
        for e in range(envs):
            seq = [None if acts_log[k][e]["shape"] < 0 else
                   tuple(int(acts_log[k][e][f]) for f in ("target_block", "target_face", "shape", "face")) +
                   (float(acts_log[k][e]["offset_x"]), float(acts_log[k][e]["offset_y"])) for k in range(steps)]
            jobs.append(dict(shapes=spec["shapes"], obstacles=spec["obstacles"], targets=spec["targets"], mu=0.8,
                             max_steps=wargs.max_steps, actions=seq, x_ground=X_GROUND, offsets=(0.0,)))
        traces = H.replay_parallel(jobs)
        c = dict(records=0, raster_mismatches=0, interface_mismatches=0, verdict_mismatches=0, reward_mismatches=0,
                 termination_mismatches=0, distance_mismatches=0, in_band=0)
        for e in range(envs):
            for k in range(steps):
                ref, o = traces[e][k], out_log[k][e]
                if ref.get("skipped"):
                    continue
                c["records"] += 1
                c["raster_mismatches"] += [int(v) for v in bits_log[k][e]] != ref["bits"]
                c["interface_mismatches"] += int(o["n_interfaces"]) != ref["n_interfaces"]
                fr_band = ref["r_frozen"] is not None and band[0] < ref["r_frozen"] < band[1]
                un_band = ref["r_unfrozen"] is not None and band[0] < ref["r_unfrozen"] < band[1]
                c["in_band"] += int(fr_band) + int(un_band)
                if not fr_band:
                    c["verdict_mismatches"] += bool(o["stable"]) != ref["frozen"]
                    c["reward_mismatches"] += float(o["reward"]) != ref["reward"]
                    c["termination_mismatches"] += (bool(o["terminated"]) != ref["terminated"]) or (bool(o["truncated"]) != ref["truncated"])
                if not un_band:
                    c["verdict_mismatches"] += bool(o["stable_unfrozen"]) != ref["stable_unfrozen"]
                c["distance_mismatches"] += list(o["distance_to_targets"][:len(ref["distance"])]) != ref["distance"]
        c = {k: int(v) for k, v in c.items()}
        c["ok"] = all(v == 0 for k, v in c.items() if k.endswith("mismatches")) and c["records"] > 0
        out["workloads"][name] = c
    out["ok"] = all(w["ok"] for w in out["workloads"].values())
    out["seconds"] = time.perf_counter() - t0
    out["checker"] = "oracle/ (CPU restatement of the reference) on %d host processes" % min(envs, os.cpu_count() or 1)
    return out


def run_batch_scan(job, wargs, sizes):
    """The same step at larger lock-step batches (device-timed, L2 flushed between steps): at 1024 envs a launch is
    a partial wave whose length is the slowest environment's solve; with more environments per GPU the SMs fill up
    and the per-step cost approaches the mean solve."""
    import numpy as np
    torch = job.torch
    from bridges_b200.envs.batched import BatchedAssemblyGym
    args, dev = job.args, job.dev
    spec = task_spec(wargs)
    rows = []
    for E in sizes:
        env = BatchedAssemblyGym(E, [f"shapes/{nm}.urdf" for nm in spec["shapes"]], max_steps=wargs.max_steps, device=job.local_rank)
        env.reset(dict(obstacles=spec["obstacles"], targets=spec["targets"]))
        amax = min(1024, max(128, 64 * ((env.max_candidates(len(X_GROUND), 1) + 63) // 64)))
        img = torch.empty((E, 1, 64, 64), dtype=torch.float32, device=dev)
        binary = torch.empty((E, 6), dtype=torch.float32, device=dev)
        K, W0 = 60, 3 * wargs.max_steps
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        n_pre = []
        for i in range(W0 + K):
            env.enumerate_actions(X_GROUND, (0.0,), amax=amax, with_bits=False)
            acts = env.select_random(seed=args.seed * 1000003 + i * 7919 + 17)[0]
            if i >= W0:
                job.flush.fill_(i & 0xff)
                ev[i - W0][0].record()
            env.step(acts, block_img=img, binary=binary)
            if i >= W0:
                ev[i - W0][1].record()
                if i % 20 == 0:
                    n_pre.append(np.maximum(env.read_out()["n_blocks"].astype(np.float64) - 1, 0).mean())
            env.reset_done()
        torch.cuda.synchronize()
        t = sum(a.elapsed_time(b) for a, b in ev) * 1e-3
        n_ot = len(spec["obstacles"]) + len(spec["targets"])
        bytes_per_launch = E * (16528 + 16 * n_ot + 32 * float(np.mean(n_pre)))
        rows.append({"envs": E, "env_steps_per_s": E * K / t, "ms_per_step": 1e3 * t / K,
                     "achieved_GBps": bytes_per_launch * K / t / 1e9, "hbm_frac": bytes_per_launch * K / t / 1e9 / job.peak})
        env.close()
        del env, img, binary
    return rows


def run_gpu(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist

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

    single = args.workload != "north_star"
    head_name = args.workload if single else HEADLINE
    head_args = args if args.workload == "custom" else workload_args(args, head_name)

    # CPU baselines first (rank 0, N=1 only), before this process touches CUDA in earnest
    cpu_base, cpu_secondary = None, {}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_base = cpu_baseline_record(head_args, args.cpu_budget)
        if not single and not args.no_secondary:
            for name in SECONDARY:
                cpu_secondary[name] = cpu_baseline_record(workload_args(args, name), min(6.0, args.cpu_budget), with_one_core=False)

    job = Job(args, rank, local_rank, world)
    K, W = args.steps, args.warmup
    head = measure_workload(job, head_args, K, W, args.steady_seconds)
    secondary = {}
    if not single and not args.no_secondary:
        for name in SECONDARY:
            secondary[name] = measure_workload(job, workload_args(args, name), K, W, min(args.steady_seconds, 1.0),
                                               with_e2e_variants=False)
    sweep = run_sweep(job) if (args.sweep or (not single and not args.no_sweep)) else None
    rollout = run_rollout(job, head_args) if (not args.no_rollout) else None
    scan = run_batch_scan(job, head_args, [4096, 16384, 65536]) if (args.batch_scan and rank == 0) else None
    gate = None
    if not args.no_parity_gate:
        gate = run_parity_gate(job, [head_name] if (single and head_name in WORKLOADS) else ([] if single else [HEADLINE, "tower2"]))
    job.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    line = {"metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic"}
    for k in ("config", "clocks", "e2e", "gpu_launches", "roofline", "roofline_fp64", "env_stats", "with_candidate_stage",
              "steady_state", "wall_s_timed_region"):
        if k in head:
            line[k] = head[k]
    line["e2e"]["host_binding"] = (f"rank bound to the {len(numa_cpus)} CPUs next to its GPU" if numa_cpus else "none")
    if secondary:
        for name, rec in secondary.items():
            if name in cpu_secondary:
                rec["cpu_baseline"] = cpu_secondary[name]
        line["secondary"] = {"note": "the same measurement on the other env-loop configurations of BASELINE.json",
                             "configs[1]": secondary.get("tower2"), "configs[2]": secondary.get("tower4")}
    if sweep is not None:
        line["sweep"] = sweep
    if rollout is not None:
        line["rollout"] = rollout
    if scan is not None:
        line["batch_scan"] = {"note": "same step kernel and workload at larger lock-step batches per GPU (rank 0)", "rows": scan}
    if gate is not None:
        line["parity_gate"] = gate
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
    ap.add_argument("--workload", default=None, choices=["north_star", "bridge", "tower2", "tower4"],
                    help="north_star (default): headline on bridge(5) with the mixed library + secondary records on "
                         "configs[1] / configs[2] + the 65,536-assembly sweep + rollout + parity gate; a name: only that "
                         "env-loop workload")
    ap.add_argument("--task", default=None, choices=["tower", "bridge"],
                    help="custom single workload (with --tower-height / --num-obstacles / --shapes / --max-steps)")
    ap.add_argument("--tower-height", type=int, default=2)
    ap.add_argument("--num-obstacles", type=int, default=5, help="--task bridge: obstacle cubes to span")
    ap.add_argument("--shapes", default="trapezoid,hexagon", help="--task bridge: block library (names under shapes/)")
    ap.add_argument("--max-steps", type=int, default=None)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--cpu-budget", type=float, default=15.0, help="seconds of CPU-baseline sampling")
    ap.add_argument("--e2e-steps", type=int, default=200)
    ap.add_argument("--steady-seconds", type=float, default=1.5, help="device seconds of the steady-state sample (0 = off)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-cores", type=int, default=0, help="CPU-arm worker processes (0 = all host cores)")
    ap.add_argument("--no-numa-bind", action="store_true", help="N>1: do not bind the rank to the CPUs of its GPU's NUMA node")
    ap.add_argument("--no-flush", action="store_true", help="profiling only: skip the L2 flush between steps")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--no-sweep", action="store_true")
    ap.add_argument("--no-rollout", action="store_true")
    ap.add_argument("--no-parity-gate", action="store_true")
    ap.add_argument("--sweep", action="store_true", help="single-workload runs: also run the 65,536-assembly sweep")
    ap.add_argument("--batch-scan", action="store_true", help="also time the step at 4096 / 16384 / 65536 envs per GPU")
    ap.add_argument("--sweep-assemblies", type=int, default=65536)
    ap.add_argument("--sweep-steps", type=int, default=10)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.warmup < 3:
        args.warmup = 3
    if args.task is not None:                      # explicit custom workload
        args.workload = "custom"
        if args.max_steps is None:
            args.max_steps = 15 if args.task == "bridge" else 10
    else:
        args.workload = args.workload or "north_star"
        name = HEADLINE if args.workload == "north_star" else args.workload
        for k, v in WORKLOADS[name].items():       # the reference arm and config_dict read these
            if k != "max_steps" or args.max_steps is None:
                setattr(args, k, v)
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_gpu(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
