"""act() path (configs[2]): encoder + actor, quadruped shape (A=12): latency at batch 1 (what train.py calls
every environment step) and throughput at batch 1024 (vectorised rollout), both modes.  Host numpy in,
host numpy out, as the reference's act (drqv2.py:164-175)."""
import json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from drqv2_b200 import DrQV2Agent

out = {}
for mode in ("bf16", "fp32"):
    agent = DrQV2Agent((9, 84, 84), (12,), "cuda", 1e-4, 50, 1024, 0.01, 2000, 2, bench.SCHED, 0.3, False, seed=0, mode=mode)
    o1 = np.random.randint(0, 256, (9, 84, 84), dtype=np.uint8)
    ob = np.random.randint(0, 256, (1024, 9, 84, 84), dtype=np.uint8)
    for obs, tag, n in ((o1, "b1", 500), (ob, "b1024", 30)):
        for _ in range(5):
            agent.act(obs, 5000, False)
        torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(n):
            a = agent.act(obs, 5000, False)
        dt = (time.perf_counter() - t) / n
        out[f"{mode}_{tag}_ms"] = dt * 1e3
        if tag == "b1024":
            out[f"{mode}_{tag}_obs_per_s"] = 1024 / dt
        assert np.all(np.abs(a) <= 1.0)
print(json.dumps(out))
