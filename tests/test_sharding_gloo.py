"""The N > 1 host logic on CPU with the gloo backend (world_size 2): contiguous env sharding,
max-over-ranks timing, gather of per-shard verdicts, and the rank-0-only reference arm."""
import json
import os
import subprocess
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    from bridges_b200.sharding import gather_to_rank0, max_over_ranks, shard_range
    dist.init_process_group("gloo", rank=rank, world_size=world)
    total = 65537                                    # deliberately not divisible
    lo, hi = shard_range(total, rank, world)
    local = torch.arange(lo, hi, dtype=torch.int32) % 251
    t = max_over_ranks([1.0 + rank, 5.0 - rank])
    allv = gather_to_rank0(local, total, rank, world)
    if rank == 0:
        ok = bool(torch.equal(allv, torch.arange(total, dtype=torch.int32) % 251))
        json.dump(dict(t=t, ok=ok, lo=lo, hi=hi), open(os.path.join(tmp, "r0.json"), "w"))
    else:
        json.dump(dict(t=t, lo=lo, hi=hi), open(os.path.join(tmp, "r1.json"), "w"))
    dist.destroy_process_group()


def test_shard_and_reduce_world2(tmp_path):
    mp.spawn(_worker, args=(2, 29611, str(tmp_path)), nprocs=2, join=True)
    r0 = json.load(open(tmp_path / "r0.json"))
    r1 = json.load(open(tmp_path / "r1.json"))
    assert r0["ok"]
    assert r0["t"] == r1["t"] == [2.0, 5.0]
    assert (r0["lo"], r0["hi"], r1["lo"], r1["hi"]) == (0, 32768, 32768, 65537)


def test_shard_range_partitions():
    from bridges_b200.sharding import shard_range
    for total in (1, 7, 1024, 65536):
        for world in (1, 2, 4, 8):
            edges = [shard_range(total, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == total
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))


def test_reference_arm_prints_on_rank0_only():
    env = dict(os.environ, WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29612")
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
           "--warmup", "3", "--cpu-cores", "2"]
    r1 = subprocess.run(cmd, env=dict(env, RANK="1", LOCAL_RANK="1"), capture_output=True, text=True, timeout=300)
    assert r1.returncode == 0 and r1.stdout.strip() == ""
    r0 = subprocess.run(cmd, env=dict(env, RANK="0", LOCAL_RANK="0"), capture_output=True, text=True, timeout=300)
    assert r0.returncode == 0, r0.stderr
    line = json.loads(r0.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["n_gpus"] == 2 and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0
    assert line["unit"] == "env_steps/s" and line["higher_is_better"] is True


def test_numa_binding_helpers():
    """`bind_to_gpu_numa_node` is best effort: a parsed cpulist, and no change without a CUDA device."""
    import os
    from bridges_b200.sharding import _parse_cpulist, bind_to_gpu_numa_node
    assert _parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert _parse_cpulist("") == set()
    before = os.sched_getaffinity(0)
    import torch
    if not torch.cuda.is_available():
        assert bind_to_gpu_numa_node(0) is None
        assert os.sched_getaffinity(0) == before
