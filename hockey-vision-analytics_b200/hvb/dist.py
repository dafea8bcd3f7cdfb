"""Multi-GPU plumbing (one process per GPU, torch.distributed).

The per-frame path shards with NO data-path collective: clips / frame chunks are partitioned across
ranks (`shard_range`) and each rank runs the same kernels on its shard.  The ONE exchange step is
collecting crop features for the global team-clustering fit: a count all-gather
followed by a padded all-gather over NCCL (NVLink 5 / NVSwitch), compacted in rank order so every
rank holds the bit-identical [sum N_g, D] matrix and computes identical scaler statistics.
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of n_items owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_clips(n_clips: int, rank: int, world: int) -> List[int]:
    """Clip indices owned by `rank`: stateful stages (ByteTrack, temporal vote) stay with the owner."""
    lo, hi = shard_range(n_clips, rank, world)
    return list(range(lo, hi))


def all_gather_features(local: torch.Tensor, group=None) -> torch.Tensor:
    """local: [N_g, D] (any float dtype, N_g may be 0 and differ per rank) -> [sum N_g, D], rank order."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    d = local.shape[1]
    n_local = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local, group=group)
    counts = [int(c.item()) for c in counts]
    n_max = max(counts)
    if n_max == 0:
        return local
    padded = torch.zeros((n_max, d), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    gathered = torch.empty((world * n_max, d), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(gathered, padded.contiguous(), group=group)
    parts = [gathered[r * n_max: r * n_max + counts[r]] for r in range(world)]
    return torch.cat(parts, 0).contiguous()
