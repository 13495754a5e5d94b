"""Pair sharding for multi-GPU inference: pairs are independent, so a batch is split
into contiguous blocks, one per rank, balanced by the attention-dominated cost model; no collective
runs on the hot path, results are gathered afterwards."""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def pair_cost(n0: int, n1: int) -> float:
    return float(n0 * n0 + n1 * n1 + 1.5 * n0 * n1)


def shard_bounds(costs: Sequence[float], world: int) -> List[Tuple[int, int]]:
    """Contiguous [start, end) per rank; every pair assigned exactly once; prefix balancing by cost.
    When there are at least `world` pairs every rank gets at least one."""
    n = len(costs)
    total = float(sum(costs))
    bounds, start, acc = [], 0, 0.0
    for r in range(world):
        remaining_ranks = world - r - 1
        if r == world - 1:
            end = n
        else:
            target = total * (r + 1) / world
            end = start
            while end < n - remaining_ranks and (end == start or acc + costs[end] <= target + 1e-9):
                acc += costs[end]
                end += 1
            if n < world:
                end = min(start + (1 if start < n else 0), n)
        bounds.append((start, end))
        start = end
    return bounds


def gather_to_rank0(t: torch.Tensor, sizes: Sequence[int], group=None) -> Optional[torch.Tensor]:
    """Concatenates per-rank [b_r, ...] tensors on rank 0 (b_r may differ per rank)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return t
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    mx = max(sizes)
    pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    out = [torch.empty_like(pad) for _ in range(world)] if rank == 0 else None
    dist.gather(pad, out, dst=0, group=group)
    if rank != 0:
        return None
    return torch.cat([o[:s] for o, s in zip(out, sizes)], 0)
