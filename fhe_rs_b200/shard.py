"""Sharding of independent ciphertexts across ranks (one process per GPU).

Every ciphertext operation of the path reads only its operands and the immutable, replicated parameter /
key tables (bfv/ops/mul.rs:165), so the batch dimension is split into contiguous blocks per rank and there is
NO data-path collective.  torch.distributed (NCCL on GPUs, gloo in the CPU tests) is used only for the
barrier, the max-over-ranks timing and the optional gather of result checksums."""
from __future__ import annotations

from typing import List, Tuple


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """[first, last) block of `total` items owned by `rank`: sizes differ by at most one, order preserved."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(total, world)
    first = rank * base + min(rank, rem)
    return first, first + base + (1 if rank < rem else 0)


def max_over_ranks(value: float, device=None) -> float:
    """max of a per-rank scalar (device time of the slowest rank)"""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_checksums(local: List[int], device=None) -> List[int]:
    """all-gather of per-ciphertext 63-bit checksums, in global ciphertext order (uneven shards allowed)"""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return list(local)
    world = dist.get_world_size()
    n = torch.tensor([len(local)], dtype=torch.int64, device=device)
    sizes = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(sizes, n)
    m = max(int(s.item()) for s in sizes)
    buf = torch.zeros(m, dtype=torch.int64, device=device)
    buf[: len(local)] = torch.tensor(local, dtype=torch.int64, device=device)
    out = [torch.zeros(m, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(out, buf)
    res: List[int] = []
    for s, o in zip(sizes, out):
        res += [int(x) for x in o[: int(s.item())].cpu()]
    return res


def bind_host_thread_to_gpu(device_index: int) -> str:
    """Pin the calling host thread (and thereby the first-touch placement of the pinned staging buffers it allocates
    next) to the CPUs of the NUMA node the GPU hangs off.  On two-socket boxes a staging buffer on the far socket
    halves the PCIe rate of the end-to-end path.  Returns a short description of what was done; never raises."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        props = torch.cuda.get_device_properties(device_index)
        handle = None
        uuid = getattr(props, "uuid", None)
        if uuid is not None:
            try:
                handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(uuid)).encode())
            except Exception:
                handle = None
        if handle is None and hasattr(props, "pci_bus_id"):
            bus = "%08x:%02x:%02x.0" % (getattr(props, "pci_domain_id", 0), props.pci_bus_id, props.pci_device_id)
            handle = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        if handle is None:
            return "no NVML handle"
        pynvml.nvmlDeviceSetCpuAffinity(handle)
        import os
        return "bound to %d CPUs of the GPU's NUMA node" % len(os.sched_getaffinity(0))
    except Exception as e:  # noqa: BLE001 -- affinity is an optimisation, not a requirement
        return "not bound (%s)" % type(e).__name__
