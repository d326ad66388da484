"""The data-parallel CUDA path where only ONE GPU is available (the driver's GPU test box): two ranks share
cuda:0 and exchange gradients through gloo (staged through the host, eager launches - NCCL refuses two ranks on
one device and cannot be captured then).  What is checked is the required DP property:

    DP(N ranks, each on its slice of the global batch and of the global draws)  ==  single GPU on the whole batch

for every gradient, the averaged metrics, and - with the real learning rate - bit-identical replicas.  The NCCL /
CUDA-graph variant of the same schedule runs in tests/test_gpu_dp.py when two GPUs are present."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
SCHED = "linear(1.0,0.1,100000)"


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, mode, out):
    import faulthandler
    import torch.distributed as dist
    faulthandler.dump_traceback_later(240, exit=True)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from drqv2_b200 import DrQV2Agent, dist as D
        from oracle import drq_oracle as O
        from tests.helpers import rel_l2
        A, Fd, H, Bg = 6, 50, 128, 16
        Bs = Bg // world
        params = O.synthetic_params(9, A, Fd, H, seed=4)
        keys = ("obs", "action", "reward", "discount", "next_obs")
        draws = ("shift_obs", "shift_next", "eps_critic", "eps_actor")

        def make(dp, lr):
            ag = DrQV2Agent((9, 84, 84), (A,), "cuda", lr, Fd, H, 0.01, 2000, 2, SCHED, 0.3, True,
                            use_cuda_graph=False, seed=5, mode=mode, data_parallel=dp)
            for net in ("encoder", "actor", "critic", "critic_target"):
                getattr(ag, net).load_state_dict(params[net])
            return ag

        def run(ag, b, step, sl=slice(None)):
            ag.inject_draws(*[b[k][sl] for k in draws])
            return ag.update(iter([tuple(b[k][sl] for k in keys)]), step)

        sl = slice(rank * Bs, (rank + 1) * Bs)
        b = O.synthetic_batch(Bg, A, seed=100)
        # ---- lr = 0: every gradient of the sharded update equals the whole-batch update's
        dp0, solo0 = make(True, 0.0), make(False, 0.0)
        assert dp0.data_parallel and not solo0.data_parallel
        m_dp, m_solo = run(dp0, b, 0, sl), run(solo0, b, 0)
        torch.cuda.synchronize()
        worst = 0.0
        for net in ("encoder", "critic", "actor"):
            for (name, p), (_, q) in zip(getattr(dp0, net).named_parameters(), getattr(solo0, net).named_parameters()):
                worst = max(worst, rel_l2(p.grad.cpu().numpy(), q.grad.cpu().numpy()))
        for k in m_solo:                                         # the 8 metrics are averaged over the ranks
            assert abs(m_dp[k] - m_solo[k]) <= 2e-5 * abs(m_solo[k]) + 1e-6, (k, m_dp[k], m_solo[k])
        # ---- real lr, three updates: replicas stay bit-identical
        dp = make(True, 1e-4)
        for s in range(3):
            run(dp, O.synthetic_batch(Bg, A, seed=200 + s), 2 * s, sl)
        torch.cuda.synchronize()
        a = dp._arena
        same = D.replicas_identical([a.params, a.target, a.exp_avg, a.exp_avg_sq])
        moved = not torch.equal(a.params, make(False, 1e-4)._arena.params)
        out[rank] = (worst, bool(same), bool(moved))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["bf16", "fp32"])
def test_dp_two_ranks_on_one_gpu_equal_whole_batch(mode):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), mode, out), nprocs=world, join=True)
    assert len(out) == world
    for r in range(world):
        worst, same, moved = out[r]
        # per-sample arithmetic is identical; only the order of the batch sums differs (fp32 accumulation)
        assert worst <= (2e-3 if mode == "bf16" else 1e-4), dict(out)
        assert same and moved, dict(out)
