"""Image-wise sharding of the sampling core across the GPUs of one box.

The path never mixes batch entries (ms_deform_attn.py:159-193 treats every ``n``
independently), so inference shards by splitting the images over the ranks with
no data-path collective; a training step only adds the
gradient all-reduce that stock DDP already performs
(/root/reference/src/misc/dist_utils.py:126).  The helpers here do the
bookkeeping: who owns which images, how many units the whole job processed, and
an optional gather of per-rank results.  They work with any initialised
``torch.distributed`` backend (``nccl`` on the box, ``gloo`` in the CPU tests).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist

__all__ = ["image_range", "local_images", "shard_batch", "job_total", "max_over_ranks", "gather_rows", "bind_to_gpu_numa"]


def image_range(n_images: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Half-open range of images owned by ``rank``; the first ``n % world`` ranks take one extra."""
    if world_size <= 0 or not 0 <= rank < world_size:
        raise ValueError(f"bad rank {rank} / world_size {world_size}")
    base, extra = divmod(int(n_images), world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def local_images(n_images: int, rank: int, world_size: int) -> int:
    start, stop = image_range(n_images, rank, world_size)
    return stop - start


def shard_batch(tensors: Sequence[torch.Tensor], rank: int, world_size: int) -> List[torch.Tensor]:
    """Slice every tensor along dim 0 (the image axis) to this rank's images."""
    n = tensors[0].shape[0]
    start, stop = image_range(n, rank, world_size)
    return [t[start:stop] for t in tensors]


def _world(group=None) -> int:
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def job_total(local_units: float, device=None, group=None) -> float:
    """Sum of a per-rank count over the job (units all ranks processed)."""
    if _world(group) == 1:
        return float(local_units)
    t = torch.tensor([float(local_units)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return float(t.item())


def max_over_ranks(value: float, device=None, group=None) -> float:
    """Max of a per-rank measurement (device time of a step) over the job."""
    if _world(group) == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def gather_rows(local: torch.Tensor, n_images: int, group=None) -> torch.Tensor:
    """All-gather per-rank result rows (dim 0 = this rank's images) back into image order."""
    world = _world(group)
    if world == 1:
        return local
    rank = dist.get_rank(group)
    counts = [local_images(n_images, r, world) for r in range(world)]
    if local.shape[0] != counts[rank]:
        raise ValueError(f"rank {rank} holds {local.shape[0]} rows, expected {counts[rank]}")
    width = max(counts)
    padded = local.new_zeros((width,) + tuple(local.shape[1:]))
    padded[:local.shape[0]] = local
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)


def bind_to_gpu_numa(device_index: int) -> str:
    """Pin the calling process to the CPU cores NVML reports as local to GPU ``device_index`` (one process
    per GPU): pinned host buffers are then first-touched on the GPU's own NUMA node, which is what keeps the
    host<->device copies of N ranks from sharing one memory controller.  Returns a short description;
    never raises (no NVML, no permission: the process keeps its affinity)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        index = device_index
        if visible:
            entry = visible.split(",")[device_index].strip()
            if entry.isdigit():
                index = int(entry)
        handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        cpus = [w * 64 + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return "unchanged (no overlap with the allowed CPUs)"
        os.sched_setaffinity(0, allowed)
        return f"nvml-local cores of GPU {index} ({len(allowed)} cpus)"
    except Exception as exc:                      # noqa: BLE001 - diagnostics only
        return f"unchanged ({type(exc).__name__})"
