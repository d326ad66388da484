"""Driver for ncu / timing of the fused optimiser launches alone.  usage: run_opt.py [iters]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from drqv2_b200 import DrQV2Agent

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3
agent = DrQV2Agent((9, 84, 84), (6,), "cuda", 1e-4, 50, 1024, 0.01, 2000, 2, bench.SCHED, 0.3, False, use_cuda_graph=False, seed=0, mode="bf16")
st, a = agent._bf16, agent._arena
a.grads.normal_(0, 1e-3)
for o in (0, 16, 24):
    agent._scal_dev[o:o + 6] = torch.tensor([0.1, 0.999, 0.001, 0.05, 1e-8, -1e-4])
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for f in (st.step_critic_encoder, st.step_actor_target):
    ts = []
    for i in range(iters):
        flush.fill_(i)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    print(f.__name__, "us (L2 flushed):", [round(t, 1) for t in ts])
