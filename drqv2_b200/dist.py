"""Multi-GPU plumbing of the update: one process per GPU, torch.distributed only.

* ensembles of independent agents need nothing from here (no collective on the data path);
* large-batch data-parallel updates shard the batch over ranks, replicate parameters and optimiser
  state, and average the gradient arenas between the backward and the optimiser kernels: one
  all-reduce over the contiguous [encoder | critic] gradient range after the critic backward and one
  over the actor range after the actor backward.  Losses are batch means (drqv2.py:189,216) and there is
  no batch-norm, so the mean over ranks of the per-shard mean gradients is the full-batch gradient.

The collective runs on the caller's current CUDA stream, so it is captured into the update's CUDA graph
(NCCL) - or runs on CPU tensors with gloo in the tests.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def world(group=None):
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def shard_sizes(global_batch, n):
    """Rows of the global batch owned by each rank (equal shards are required for mean-of-means)."""
    if global_batch % n != 0:
        raise ValueError(f"global batch {global_batch} does not split evenly over {n} ranks")
    return [global_batch // n] * n


def rank_seed(seed, rank):
    """Device RNG key of a rank: every shard draws its own shifts and noise."""
    return (int(seed) * 0x9E3779B97F4A7C15 + int(rank) * 0xBF58476D1CE4E5B9 + 1) & 0x7FFFFFFFFFFFFFFF


def average_(t, group=None):
    """In-place mean over ranks of a flat gradient range (NCCL: one AVG all-reduce; gloo: SUM then scale)."""
    n = world(group)
    if n == 1:
        return t
    if t.is_cuda:
        dist.all_reduce(t, op=dist.ReduceOp.AVG, group=group)
    else:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        t.div_(n)
    return t


def broadcast_(tensors, src=0, group=None):
    """Make parameter / optimiser-state arenas identical on every rank (construction time)."""
    if world(group) == 1:
        return
    for t in tensors:
        dist.broadcast(t, src=src, group=group)
