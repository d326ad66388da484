"""Multi-GPU plumbing of the update: one process per GPU, torch.distributed only.

* ensembles of independent agents need nothing from here (no collective on the data path);
* large-batch data-parallel updates shard the batch over ranks, replicate parameters and optimiser
  state, and average the gradient arenas between the backward and the optimiser kernels.  Losses are batch
  means (drqv2.py:189,216) and there is no batch-norm, so the mean over ranks of the per-shard mean gradients
  is the full-batch gradient.  Schedule of the tensor-core mode (drqv2_b200/_bf16.py, DrQV2Agent._update_body):
    - the critic's gradients (99.5 % of the [encoder | critic] bytes) are complete before the encoder backward
      starts; their all-reduce, critic_opt.step() and the whole actor pass run on the main stream BESIDE the
      encoder backward on the side stream - the encoder is dead after drqv2.py:246;
    - the actor's gradients and the update's 8 metrics (stored right behind them) share one all-reduce;
    - the encoder's 30 k gradient floats are averaged after the streams join, then encoder_opt.step().
  All collectives are issued from the main stream in one fixed order on every rank (one communicator must not
  run two collectives concurrently), which is why the encoder's small all-reduce waits for the join instead
  of riding the side stream.  The fp32 parity mode keeps one stream: [encoder | critic] in one all-reduce.

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
    """In-place mean over ranks of a flat gradient range (NCCL: one AVG all-reduce, graph-capturable; gloo - CPU
    tensors, or CUDA tensors staged through the host in the tests - SUM then scale)."""
    n = world(group)
    if n == 1:
        return t
    if dist.get_backend(group) == "nccl":
        dist.all_reduce(t, op=dist.ReduceOp.AVG, group=group)
    else:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        t.div_(n)
    return t


def replicas_identical(tensors, group=None):
    """True when the given tensors hold the same bits on every rank (data-parallel replicas must: they start from
    a broadcast and apply the same averaged gradients).  Compares an exact 64-bit checksum per tensor."""
    n = world(group)
    if n == 1:
        return True
    sums = []
    for t in tensors:
        bits = t.detach().contiguous().view(torch.int32).to(torch.int64)
        idx = torch.arange(1, bits.numel() + 1, device=bits.device, dtype=torch.int64)
        sums.append((bits * (idx % 8191 + 1)).sum())          # position-weighted: permutations change it
    mine = torch.stack(sums)
    if dist.get_backend(group) != "nccl":
        mine = mine.cpu()
    every = [torch.zeros_like(mine) for _ in range(n)]
    dist.all_gather(every, mine, group=group)
    return all(torch.equal(every[0], e) for e in every[1:])


def broadcast_(tensors, src=0, group=None):
    """Make parameter / optimiser-state arenas identical on every rank (construction time)."""
    if world(group) == 1:
        return
    for t in tensors:
        dist.broadcast(t, src=src, group=group)
