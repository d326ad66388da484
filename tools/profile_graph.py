"""Driver for `ncu --graph-profiling node --cache-control none`: a few CUDA-graph replays of the bench
workload so that every kernel node is timed with warm caches.  usage: profile_graph.py [replays] [mode]"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from drqv2_b200 import DrQV2Agent, make_replay_loader

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
mode = sys.argv[2] if len(sys.argv) > 2 else "bf16"
B, A, Fd, H = 256, 6, 50, 1024
torch.manual_seed(0); np.random.seed(7)
agent = DrQV2Agent((9, 84, 84), (A,), "cuda", 1e-4, Fd, H, 0.01, 2000, 2, bench.SCHED, 0.3, False, use_cuda_graph=True, seed=0, mode=mode)
bench.fill_ring("/prof/ring", A, 16, 501, torch.device("cuda"))
it = iter(make_replay_loader("/prof/ring", 16 * 501, B, 0, False, 3, 0.99))
for i in range(3 + n):
    agent.update(it, 2 * i)
torch.cuda.synchronize()
print("done", n, "replays")
