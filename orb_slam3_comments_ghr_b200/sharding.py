"""Query sharding for the batched workloads: frame-pair batches (C4) and kNN
queries (C5) are split contiguously by row over the ranks of one box; each rank works on its rows
with no data-path collective, and ONE all-gather of the match indices assembles the result.
The database / keyframe set is replicated.  Works with NCCL (GPU tensors) and gloo (CPU tests)."""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """rows [lo, hi) of rank: contiguous, sizes differ by at most one, earlier ranks get the extra row"""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def max_shard(n: int, world: int) -> int:
    return (n + world - 1) // world


def all_gather_rows(local: torch.Tensor, n_total: int, out: torch.Tensor = None) -> torch.Tensor:
    """local: [rows_of_this_rank, ...] -> [n_total, ...] on every rank with a single all-gather.
    Equal shards (n_total divisible by the world size) gather straight into `out` (a caller-owned [n_total, ...]
    buffer, allocated when omitted) with no staging copy; ragged shards are padded to the largest shard and trimmed."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    cap = max_shard(n_total, world)
    if n_total % world == 0:
        if out is None:
            out = torch.empty((n_total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous())
        return out
    pad = torch.zeros((cap,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * cap,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad)
    parts = []
    for r in range(world):
        lo, hi = shard_bounds(n_total, r, world)
        parts.append(out[r * cap: r * cap + (hi - lo)])
    return torch.cat(parts, dim=0)
