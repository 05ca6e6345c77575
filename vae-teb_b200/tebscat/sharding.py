"""Batch sharding across ranks: every signal is independent
(core/scattering1d.py:269-399 has no cross-batch op), so each rank transforms a contiguous
slice of the batch with its own plan replica and there is no collective on the transform.
``torch.distributed`` is used only for barriers and for reducing timings."""
import os

import torch
import torch.distributed as dist


def bind_host_to_device(device_index: int):
    """Restrict the calling process to the CPUs NVML reports as local to the GPU (its NUMA node), so that the pinned
    staging buffers allocated afterwards are node-local: with one rank per GPU every rank streams ~30 GB/s of host
    memory through its own root complex instead of across the socket interconnect.  Returns the previous affinity
    (to restore with os.sched_setaffinity) or None when NVML / the affinity call is unavailable."""
    try:
        import pynvml
        before = os.sched_getaffinity(0)
        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(device_index)
        bus = '%08x:%02x:%02x.0' % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        handle = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        pynvml.nvmlDeviceSetCpuAffinity(handle)
        return before
    except Exception:
        return None


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous, balanced slice [lo, hi) of n_items for `rank` of `world` (sizes differ by <= 1)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError('rank {} outside world of {}'.format(rank, world))
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def max_over_ranks(value: float, device=None) -> float:
    """Max of a per-rank scalar (timings are reported as the slowest rank)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sharded_apply(fn, x: torch.Tensor, gather: bool = False):
    """Apply `fn` to this rank's slice of the batch dimension of x.  Returns (lo, hi, y_local) or,
    with gather=True, the full result assembled on every rank (for tests; production keeps the
    shards on their devices)."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    lo, hi = shard_range(x.shape[0], rank, world)
    y = fn(x[lo:hi])
    if not gather or world == 1:
        return lo, hi, y
    sizes = [shard_range(x.shape[0], r, world) for r in range(world)]
    parts = [torch.empty((h - l,) + tuple(y.shape[1:]), dtype=y.dtype, device=y.device) for l, h in sizes]
    dist.all_gather(parts, y.contiguous()) if len({p.shape for p in parts}) == 1 else \
        [dist.broadcast(parts[r] if r != rank else parts[r].copy_(y), src=r) for r in range(world)]
    return lo, hi, torch.cat(parts, 0)
