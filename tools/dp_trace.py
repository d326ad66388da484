"""Kernel timeline of a few data-parallel updates on rank 0 (torch profiler / CUPTI): where the time between
the ranks' collectives goes.  Run under torchrun: dp_trace.py [global_batch]"""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from drqv2_b200 import DrQV2Agent, make_replay_loader

rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
GB = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
B, A, Fd, H = GB // world, 21, 100, 1024
torch.manual_seed(0); np.random.seed(7 + rank)
agent = DrQV2Agent((9, 84, 84), (A,), "cuda", 1e-4, Fd, H, 0.01, 2000, 2, bench.SCHED, 0.3, False, use_cuda_graph=True,
                   seed=0, mode="bf16", data_parallel=True)
bench.fill_ring(f"/trace/ring{rank}", A, 16, 501, torch.device("cuda"), seed=1 + rank)
it = iter(make_replay_loader(f"/trace/ring{rank}", 16 * 501, B, 0, False, 3, 0.99))
step = 0
for _ in range(8):
    agent.update(it, step); step += 2
torch.cuda.synchronize(); dist.barrier()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(4):
        agent.update(it, step); step += 2
    torch.cuda.synchronize()
if rank == 0:
    evs = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
    t0 = evs[0].time_range.start
    # last update only
    starts = [i for i, e in enumerate(evs) if "ring_sample" in e.name]
    last = evs[starts[-1]:]
    tot = last[-1].time_range.end - last[0].time_range.start
    busy = sum(e.time_range.end - e.time_range.start for e in last)
    print(f"last update: {len(last)} kernels, span {tot:.0f} us, busy {busy:.0f} us")
    prev_end = last[0].time_range.start
    for e in last:
        d = e.time_range.end - e.time_range.start
        gap = e.time_range.start - prev_end
        if d > 60 or gap > 20 or "nccl" in e.name.lower():
            print(f"  +{e.time_range.start - last[0].time_range.start:8.0f} us  dur {d:7.0f}  gap-before {gap:6.0f}  {e.name[:70]}")
        prev_end = e.time_range.end
agent._graphs.clear()
torch.cuda.synchronize(); dist.barrier()
os._exit(0)
