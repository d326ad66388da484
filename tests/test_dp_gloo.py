"""Host-side logic of the data-parallel path (drqv2_b200/dist.py) with world_size 2 on the gloo backend
(CPU): gradient averaging, parameter broadcast, shard / seed helpers, and - with the oracle as the checker -
that the mean over ranks of the per-shard gradients of one update is the full-batch gradient
(drqv2.py:189,216 are batch means; there is no batch statistic anywhere in the networks)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from drqv2_b200 import dist as D


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        from oracle import drq_oracle as O
        # ---- helpers
        t = torch.arange(8, dtype=torch.float32) * (rank + 1)
        D.average_(t)
        assert torch.equal(t, torch.arange(8, dtype=torch.float32) * 1.5)
        p = torch.full((5,), float(rank + 3))
        D.broadcast_([p])
        assert torch.equal(p, torch.full((5,), 3.0))
        assert D.world() == world and D.shard_sizes(8, world) == [4, 4]
        assert D.rank_seed(5, 0) != D.rank_seed(5, 1)
        # ---- DP math: per-shard critic+encoder and actor gradients (lr = 0 so both passes see the same
        # parameters), averaged over ranks, equal the gradients of the full batch
        A, Fd, H, B = 4, 16, 32, 8
        params = O.synthetic_params(9, A, Fd, H, seed=3)
        b = O.synthetic_batch(B, A, seed=11)
        sched = "linear(1.0,0.1,100000)"
        keys = ("obs", "action", "reward", "discount", "next_obs")
        draws = ("shift_obs", "shift_next", "eps_critic", "eps_actor")
        sl = slice(rank * B // world, (rank + 1) * B // world)
        shard = O.OracleAgent(params, 0.0, 0.01, sched, 0.3, dtype=torch.float64)
        shard.update(*[b[k][sl] for k in keys], 0, *[b[k][sl] for k in draws])
        full = O.OracleAgent(params, 0.0, 0.01, sched, 0.3, dtype=torch.float64)
        full.update(*[b[k] for k in keys], 0, *[b[k] for k in draws])
        worst = 0.0
        for net in ("encoder", "critic", "actor"):
            for name, g in shard.grads[net].items():
                g = g.clone()
                D.average_(g)
                ref = full.grads[net][name]
                worst = max(worst, float((g - ref).abs().max() / (ref.abs().max() + 1e-30)))
        out[rank] = worst
    finally:
        dist.destroy_process_group()


def test_dp_helpers_and_mean_of_shard_gradients():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert len(out) == world
    for r in range(world):
        assert out[r] <= 1e-9, dict(out)


def test_shard_sizes_reject_ragged():
    with pytest.raises(ValueError):
        D.shard_sizes(10, 4)
