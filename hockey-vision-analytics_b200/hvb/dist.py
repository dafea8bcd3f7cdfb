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


_hvb_comms = {}


def hvb_comm(device, group=None):
    """The libhvb communicator of (device, group): created on first use — rank 0 makes the NCCL unique id
    (hvb_comm_unique_id), torch.distributed carries its 128 bytes to the other ranks (any backend), every rank calls
    hvb_comm_create.  Collective."""
    from .runtime import get_context
    ctx = get_context(device)
    key = (str(ctx.device), id(group))
    if key not in _hvb_comms:
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        box = [ctx.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        _hvb_comms[key] = (ctx, ctx.comm_create(box[0], world, rank), world)
    return _hvb_comms[key]


def all_gather_features(local: torch.Tensor, group=None, backend: str = "torch") -> torch.Tensor:
    """local: [N_g, D] (any float dtype, N_g may be 0 and differ per rank) -> [sum N_g, D], rank order.

    backend "torch": count all-gather + padded all_gather_into_tensor through torch.distributed (NCCL on GPUs, gloo in the
    CPU tests).  backend "hvb": the C ABI's own exchange (hvb_allgather_counts + hvb_allgather_features: an all-gather-v
    as one NCCL group of broadcasts into the compacted output, float64 device tensors only) — what a non-Python host
    calls; same result bit for bit."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    if backend == "hvb":
        ctx, comm, world = hvb_comm(local.device, group)
        return ctx.allgather_features(comm, local, world)[0]
    if backend != "torch":
        raise ValueError("backend must be 'torch' or 'hvb'")
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
