"""Multi-GPU layout of the inference path: images are independent end to end, so
the batch is sharded contiguously across ranks (one process per GPU), weights are replicated, and
the ONLY exchange is an all-gather of the (B/G, K, 56) detections (22.4 kB per image)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int):
    """Contiguous slice [lo, hi) of a batch of ``total`` images owned by ``rank`` (earlier ranks take
    the remainder, so shard sizes differ by at most one)."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_detections(local: torch.Tensor, group=None, total=None, out=None) -> torch.Tensor:
    """All-gather per-rank detections (b_local, K, C) into (total, K, C) on every rank.  Equal shards
    use one ``all_gather_into_tensor`` (NCCL on GPUs; gloo in the CPU tests); ragged shards are padded
    to the largest shard and trimmed."""
    world = dist.get_world_size(group)
    if world == 1:
        return local
    b = local.shape[0]
    if total is None or total == b * world:
        if out is None:
            out = local.new_empty((b * world,) + tuple(local.shape[1:]))
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    bmax = -(-total // world)
    pad = local.new_zeros((bmax,) + tuple(local.shape[1:]))
    pad[:b] = local
    buf = local.new_empty((bmax * world,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(buf, pad, group=group)
    parts = []
    for r in range(world):
        lo, hi = shard_range(total, r, world)
        parts.append(buf[r * bmax: r * bmax + (hi - lo)])
    return torch.cat(parts, 0)
