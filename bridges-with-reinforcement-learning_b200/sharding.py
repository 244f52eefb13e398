"""Env-level data parallelism: independent assemblies are partitioned contiguously over the
ranks (one process per GPU); there is no collective on the step path.  Collectives are used
only to combine timings / verdict arrays after the fact."""
import os

import torch
import torch.distributed as dist


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa_node(device_index, sysfs="/sys/bus/pci/devices"):
    """Best effort: restrict this process to the CPUs next to GPU `device_index` (its PCI device's
    `local_cpulist`), so that the pinned host buffers of the zero-copy step path are allocated on, and
    written through, the GPU's own NUMA node when several ranks share one host.  Returns the CPU set that
    was applied, or None when nothing was changed (no sysfs entry, no CUDA, affinity not permitted)."""
    try:
        bus_id = torch.cuda.get_device_properties(device_index).pci_bus_id
        dom = torch.cuda.get_device_properties(device_index).pci_domain_id
        dev = torch.cuda.get_device_properties(device_index).pci_device_id
        path = os.path.join(sysfs, "%04x:%02x:%02x.0" % (dom, bus_id, dev), "local_cpulist")
        with open(path) as fh:
            local = _parse_cpulist(fh.read())
        allowed = os.sched_getaffinity(0)
        cpus = local & allowed
        if not cpus or cpus == allowed:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except (AttributeError, OSError, ValueError, RuntimeError, AssertionError):
        return None


def shard_range(total, rank, world):
    """Contiguous partition [lo, hi) of `total` items: GPU g gets [g*N/G, (g+1)*N/G)."""
    lo = (total * rank) // world
    hi = (total * (rank + 1)) // world
    return lo, hi


def max_over_ranks(values, device=None):
    """Element-wise max of a list of floats over all ranks (timings are reported as the max)."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.cpu()]


def gather_to_rank0(local, total, rank, world, device=None):
    """Concatenate per-rank 1-D uint8/int32 tensors (shards from `shard_range`) on rank 0."""
    if world == 1:
        return local
    sizes = [shard_range(total, r, world)[1] - shard_range(total, r, world)[0] for r in range(world)]
    pad = max(sizes)
    buf = torch.zeros(pad, dtype=local.dtype, device=local.device)
    buf[:local.numel()] = local
    out = [torch.zeros(pad, dtype=local.dtype, device=local.device) for _ in range(world)]
    dist.all_gather(out, buf)
    if rank != 0:
        return None
    return torch.cat([o[:n] for o, n in zip(out, sizes)])
