"""Image sharding across ranks and the gather of the person lists (BASELINE.json north_star: "images shard
embarrassingly across the 8 GPUs of one box with an NCCL gather of the final person lists").

The path has no exchange step during compute: images are independent.  Rank r owns the contiguous
block ``shard_range(n, r, world)``; after grouping, the fixed-capacity person tensors are gathered to rank 0 in rank
order, which is image order, so an N-GPU run returns exactly what a 1-GPU run returns.  ``torch.distributed`` is
plumbing: NCCL over NVLink on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple


def shard_range(n_images: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of rank ``rank``; earlier ranks take the remainder one image each."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, rem = divmod(n_images, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_people(local: Dict[str, "torch.Tensor"], dst: int = 0, group=None) -> Optional[Dict[str, "torch.Tensor"]]:
    """Gather equally shaped per-rank result tensors to ``dst``; returns the concatenation on ``dst``, None elsewhere.

    ``local`` maps names (``n_persons [B]``, ``people_xy [B,R,J,2]``, ``people_score [B,R]`` ...) to tensors whose
    leading dimension is the rank's image count (equal on all ranks: weak scaling, or pad the last shard)."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return dict(local)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    out = {} if rank == dst else None
    for name in sorted(local):
        t = local[name]
        bufs = [torch.empty_like(t) for _ in range(world)] if rank == dst else None
        dist.gather(t, bufs, dst=dst, group=group)
        if rank == dst:
            out[name] = torch.cat(bufs, dim=0)
    return out
