"""Data-parallel update on 2 GPUs (NCCL): gradient all-reduce inside the CUDA graph keeps the replicas
bit-identical, and the averaged gradient equals the mean of the per-shard gradients."""
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
    faulthandler.dump_traceback_later(90, exit=True)        # a hung collective must not eat the GPU budget
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from drqv2_b200 import DrQV2Agent
        from oracle import drq_oracle as O
        A, Fd, H, B = 6, 50, 128, 8
        params = O.synthetic_params(9, A, Fd, H, seed=4)

        def make(dp, graph):
            ag = DrQV2Agent((9, 84, 84), (A,), "cuda", 1e-4, Fd, H, 0.01, 2000, 2, SCHED, 0.3, True,
                            use_cuda_graph=graph, seed=5, mode=mode, data_parallel=dp)
            for net in ("encoder", "actor", "critic", "critic_target"):
                getattr(ag, net).load_state_dict(params[net])
            return ag

        def run(ag, b, step):
            ag.inject_draws(b["shift_obs"], b["shift_next"], b["eps_critic"], b["eps_actor"])
            return ag.update(iter([(b["obs"], b["action"], b["reward"], b["discount"], b["next_obs"])]), step)

        dp_agent, solo = make(True, True), make(False, False)
        assert dp_agent.data_parallel
        flat = lambda ag: ag._arena.grads.clone()
        for s in range(3):                                   # warm-up, capture, replay
            b = O.synthetic_batch(B, A, seed=100 + 10 * s + rank)     # every rank its own shard
            run(dp_agent, b, 2 * s)
            if s == 0:
                run(solo, b, 0)
                g_dp, g_solo = flat(dp_agent), flat(solo)
                gathered = [torch.zeros_like(g_solo) for _ in range(world)]
                dist.all_gather(gathered, g_solo)
                want = sum(gathered) / world
                # the actor pass sees critic parameters stepped with the averaged gradient, so only the
                # [encoder | critic] range is comparable with the solo runs
                a = dp_agent._arena
                end = a.seg["critic"][0] + a.seg["critic"][2]
                err = (g_dp[:end] - want[:end]).abs().max().item() / (want[:end].abs().max().item() + 1e-30)
                assert err < 1e-5, err
        torch.cuda.synchronize()
        # replicas stay bit-identical
        mine = torch.cat([dp_agent._arena.params, dp_agent._arena.target, dp_agent._arena.grads])
        both = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(both, mine)
        assert torch.equal(both[0], both[1])
        out[rank] = True
        # captured NCCL kernels must be gone before the communicator is torn down
        dp_agent._graphs.clear()
        del dp_agent, solo
        torch.cuda.synchronize()
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["bf16", "fp32"])
def test_dp_two_gpus(mode):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 CUDA devices")
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), mode, out), nprocs=world, join=True)
    assert len(out) == world
