"""updates/s of the ring-fed graph update at the BENCH configuration, nothing else (A/B runs of an environment switch):
   DRQV2_B200_PRIO=-2,-1,0 python tools/quick_value.py [steps] [blocks]"""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from drqv2_b200 import DrQV2Agent, make_replay_loader
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
blocks = int(sys.argv[2]) if len(sys.argv) > 2 else 5
B, A, Fd, H = 256, 6, 50, 1024
torch.manual_seed(0); np.random.seed(7)
agent = DrQV2Agent((9, 84, 84), (A,), "cuda", 1e-4, Fd, H, 0.01, 2000, 2, bench.SCHED, 0.3, False, use_cuda_graph=True, seed=0, mode="bf16")
bench.fill_ring("/qv/ring", A, 16, 501, torch.device("cuda"))
it = iter(make_replay_loader("/qv/ring", 16 * 501, B, 0, False, 3, 0.99))
for i in range(10):
    agent.update(it, 2 * i)
torch.cuda.synchronize()
vals = []
for b in range(blocks):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        ws = agent.update_async(it, 2 * i)
    e1.record(); torch.cuda.synchronize()
    vals.append(steps / (e0.elapsed_time(e1) * 1e-3))
m = agent.read_metrics(ws)
print(os.environ.get("TAG", ""), "updates/s median %.1f  min %.1f max %.1f" % (sorted(vals)[len(vals) // 2], min(vals), max(vals)),
      "critic_loss", m.get("critic_loss") if isinstance(m, dict) else None, flush=True)
