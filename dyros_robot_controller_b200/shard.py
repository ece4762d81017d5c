"""Batch sharding across the GPUs of one box: every robot state is an independent problem, so rank g
of G owns the contiguous range [g*B/G, (g+1)*B/G) of the batch and there is NO collective on the solve path.  The only
communication is an optional post-solve gather of the results (<= 64 B per robot), provided here for callers that
want the full batch on every rank; it is never part of the timed control cycle.

Works with any torch.distributed backend ("nccl" on the GPUs, "gloo" in the CPU tests).
"""
from __future__ import annotations

from typing import Tuple


def shard_range(total: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous range of rank `rank`: [rank*total/world, (rank+1)*total/world) with integer floors, so the ranges
    tile [0, total) exactly and differ in length by at most one."""
    if world_size <= 0 or not (0 <= rank < world_size) or total < 0:
        raise ValueError("bad shard arguments")
    return (rank * total) // world_size, ((rank + 1) * total) // world_size


def shard_sizes(total: int, world_size: int):
    return [shard_range(total, world_size, r)[1] - shard_range(total, world_size, r)[0] for r in range(world_size)]


def gather_batch(local, total: int, group=None):
    """All-gather the per-rank result rows (a torch tensor whose first axis is this rank's shard) into batch order.
    Shards may differ in length by one row: rows are padded to the longest shard for the collective."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    sizes = shard_sizes(total, world)
    longest = max(sizes)
    if local.shape[0] != sizes[dist.get_rank(group)]:
        raise ValueError("local result does not match this rank's shard size")
    pad = torch.zeros((longest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:n] for p, n in zip(parts, sizes)], dim=0)
