"""Timeline of one graph-replayed update: kernel name, stream, start and end (us from the update's first kernel), from
torch.profiler (CUPTI).  usage: timeline.py [mode]"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from drqv2_b200 import DrQV2Agent, make_replay_loader
from torch.profiler import profile, ProfilerActivity
B, A, Fd, H = 256, 6, 50, 1024
torch.manual_seed(0); np.random.seed(7)
agent = DrQV2Agent((9, 84, 84), (A,), "cuda", 1e-4, Fd, H, 0.01, 2000, 2, bench.SCHED, 0.3, False, use_cuda_graph=True, seed=0, mode="bf16")
bench.fill_ring("/tl/ring", A, 16, 501, torch.device("cuda"))
it = iter(make_replay_loader("/tl/ring", 16 * 501, B, 0, False, 3, 0.99))
for i in range(6):
    agent.update(it, 2 * i)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(6, 10):
        agent.update_async(it, 2 * i)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start]
ev.sort(key=lambda e: e.time_range.start)
# last update: from the last prologue kernel on
starts = [i for i, e in enumerate(ev) if "prologue" in e.name]
ev = ev[starts[-1]:]
t0 = ev[0].time_range.start
streams = {}
for e in ev:
    sid = getattr(e, "device_resource_id", None) if hasattr(e, "device_resource_id") else None
    streams.setdefault(sid, len(streams))
    nm = e.name.replace("void ", "").replace("drq::", "").split("(")[0][:44]
    print(f"{e.time_range.start - t0:8.1f} {e.time_range.end - t0:8.1f} {e.time_range.end - e.time_range.start:6.1f}  s{streams[sid]}  {nm}")
print("update span", ev[-1].time_range.end - t0, "us;", len(ev), "kernels")
