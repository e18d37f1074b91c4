"""Bucket sharding across GPUs of one box: buckets are independent, so rank r of N
takes buckets r, r+N, r+2N, ... with no data-path collective.  torch.distributed is used only for the
timing barrier and the max-over-ranks reduction."""
from __future__ import annotations


def bucket_for_step(step: int, rank: int, world: int, n_buckets: int) -> int:
    """Index of the bucket that `rank` processes at (warm-up or timed) step `step`."""
    return (step * world + rank) % n_buckets


def plan(steps: int, rank: int, world: int, n_buckets: int) -> list[int]:
    return [bucket_for_step(i, rank, world, n_buckets) for i in range(steps)]


def reduce_max(values, world: int, device=None):
    """max over ranks of a list of floats (identity when world == 1)"""
    if world == 1:
        return list(values)
    import torch
    import torch.distributed as dist
    t = torch.tensor(values, dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]
