"""Multi-GPU `Y` / `L`: replicated positions, sharded term sampling.

One process per GPU (torch.distributed, NCCL over NVLink).  Terms shard naturally — every term
lives inside one path (reference src/sgd.rs:445, 502-503) — positions do not.  So:

  * rank r samples only the steps of its slice [S*r/G, S*(r+1)/G) of the concatenated step array
    and holds the records of just the paths that slice overlaps (its partners never leave them);
  * every rank keeps a full replica of the positions and runs its share
    min_term_updates * |slice| / S of every epoch, which keeps the global sampling distribution
    uniform over steps (src/sgd.rs:435, 444);
  * replicas are reconciled `syncs_per_epoch` times per epoch by an all-reduce over the position
    array: "avg" (north star: mean of the replicas) or "delta" (sum of the replicas' displacements
    since the last sync, i.e. Hogwild with staleness).

Everything here is host logic over torch tensors; it runs unchanged on CPU tensors with the gloo
backend, which is how tests/test_multi_gloo.py covers it without GPUs.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class Shard:
    rank: int
    world: int
    sample_begin: int      # global step range this rank samples from
    sample_end: int
    path_begin: int        # paths whose records this rank needs (those the slice overlaps)
    path_end: int

    @property
    def steps(self) -> int:
        return self.sample_end - self.sample_begin


def shard_steps(path_first: np.ndarray, rank: int, world: int) -> Shard:
    """Step-balanced slice for `rank` and the path range covering it."""
    path_first = np.asarray(path_first, dtype=np.uint64)
    S = int(path_first[-1])
    b = (S * rank) // world
    e = (S * (rank + 1)) // world
    if e <= b:
        return Shard(rank, world, b, b, 0, 0)
    pb = int(np.searchsorted(path_first, b, side="right") - 1)
    pe = int(np.searchsorted(path_first, e - 1, side="right"))
    return Shard(rank, world, b, e, pb, pe)


def epoch_quota(min_term_updates: int, shard: Shard, total_steps: int) -> int:
    """This rank's share of one epoch: floor/ceil split of M in proportion to the slice, exact in sum."""
    lo = (min_term_updates * shard.sample_begin) // total_steps
    hi = (min_term_updates * shard.sample_end) // total_steps
    return hi - lo


def reconcile(x, x_sync, mode: str, group=None):
    """All-reduce the replicas in place.  x: this rank's positions (torch tensor).
    "avg": x <- mean over ranks.  "delta": x <- x_sync + sum over ranks of (x - x_sync); x_sync is
    then refreshed.  Returns x."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if world == 1:
        if x_sync is not None:
            x_sync.copy_(x)
        return x
    if mode == "avg":
        dist.all_reduce(x, op=dist.ReduceOp.SUM, group=group)
        x.mul_(1.0 / world)
    elif mode == "delta":
        # sum_g x_g = G*x_sync + sum_g delta_g  =>  x_new = sum_g x_g - (G-1)*x_sync
        dist.all_reduce(x, op=dist.ReduceOp.SUM, group=group)
        x.add_(x_sync, alpha=-(world - 1))
    else:
        raise ValueError(f"unknown reconcile mode {mode!r}")
    if x_sync is not None:
        x_sync.copy_(x)
    return x
