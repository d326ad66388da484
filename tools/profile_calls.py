"""Per-entry-point device time inside one real update (eager, CUDA events around every C-ABI call,
warm caches): where the step goes, kernel by kernel.  usage: profile_calls.py [mode] [B] [iters]"""
import os
import sys
from collections import OrderedDict

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import drqv2_b200._bf16 as BF  # noqa: E402
import drqv2_b200.drqv2 as D  # noqa: E402
import drqv2_b200.replay_buffer as R  # noqa: E402
from drqv2_b200 import DrQV2Agent, _lib, make_replay_loader  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "bf16"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
A, Fd, H = 6, 50, 1024
dev = torch.device("cuda")
agent = DrQV2Agent((9, 84, 84), (A,), "cuda", 1e-4, Fd, H, 0.01, 2000, 2, bench.SCHED, 0.3, False,
                   use_cuda_graph=False, seed=0, mode=mode)
bench.fill_ring("/prof/ring", A, 64, 501, dev)
it = iter(make_replay_loader("/prof/ring", 64 * 501, B, 0, False, 3, 0.99))
for s in range(3):
    agent.update(it, 2 * s)
torch.cuda.synchronize()

records = []
orig = _lib.call


def timed(name, *a):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    orig(name, *a)
    e1.record()
    tag = name
    if name == "drq_gemm_bf16":
        tag = f"gemm M{a[12]} N{a[13]} K{a[14]} a_mn{a[2]} b_mn{a[5]} epi{a[15]} batch{a[17]} sk{a[23]} bn{a[24]}"
    elif name.startswith("drq_conv"):
        tag = f"{name} N={a[-4] if 'wgrad' not in name and 'conv1' not in name else ''}"
    records.append((tag, e0, e1))


D.call = R.call = BF.call = timed
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
for s in range(iters):
    agent.update(it, 100 + 2 * s)
t1.record()
torch.cuda.synchronize()
agg = OrderedDict()
for tag, e0, e1 in records:
    c, t = agg.get(tag, (0, 0.0))
    agg[tag] = (c + 1, t + e0.elapsed_time(e1) * 1e3)
tot = sum(t for _, t in agg.values()) / iters
print(f"mode {mode} B {B}: {len(records) // iters} calls/update, sum of call times {tot:.1f} us/update "
      f"(eager wall {t0.elapsed_time(t1) * 1e3 / iters:.1f} us)")
for tag, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t / iters:9.1f} us/update {100 * t / iters / tot:5.1f}%  x{c // iters:<3d} avg {t / c:7.1f} us  {tag}")
