"""K independent agents on one GPU, each on its own stream with its own graph and ring (BASELINE configs[3]:
ensemble members per GPU): aggregate updates/s.  usage: bench_multi_agent.py K [steps]"""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from drqv2_b200 import DrQV2Agent, make_replay_loader

K = int(sys.argv[1]) if len(sys.argv) > 1 else 2
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 300
dev = torch.device("cuda")
agents, its, streams = [], [], []
for k in range(K):
    torch.manual_seed(k)
    agents.append(DrQV2Agent((9, 84, 84), (6,), "cuda", 1e-4, 50, 1024, 0.01, 2000, 2, bench.SCHED, 0.3, False,
                             use_cuda_graph=True, seed=k, mode="bf16"))
    bench.fill_ring(f"/multi/ring{k}", 6, 16, 501, dev, seed=1 + k)
    its.append(iter(make_replay_loader(f"/multi/ring{k}", 16 * 501, 256, 0, False, 3, 0.99)))
    streams.append(torch.cuda.Stream())
step = 0
for w in range(5):
    for k in range(K):
        with torch.cuda.stream(streams[k]):
            agents[k].update(its[k], step)
    step += 2
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(steps):
    for k in range(K):
        with torch.cuda.stream(streams[k]):
            agents[k].update(its[k], step)
    step += 2
host = time.perf_counter() - t0
torch.cuda.synchronize()
wall = time.perf_counter() - t0
print(f"K={K}: {K * steps / wall:.1f} updates/s aggregate ({wall / steps * 1e3:.3f} ms per round, host issue {host / steps / K * 1e6:.1f} us per update)")
