"""GB/s of drq_adam_pack_step per segment kind (one big segment each), cold (L2 flushed) and warm."""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from drqv2_b200 import _lib
from drqv2_b200._bf16 import OptSeg, TB, OPT_PLAIN, OPT_LINEAR, OPT_TRUNK
dev = torch.device("cuda"); s = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
sc = torch.tensor([0.1, 0.999, 0.001, 0.05, 1e-8, -1e-4, 0, 0], device=dev)
def run(name, seg_fn, n, bytes_per):
    p = torch.randn(n, device=dev) * 0.1; g = torch.randn(n, device=dev) * 1e-3; m = torch.zeros(n, device=dev); v = torch.zeros(n, device=dev)
    tgt = p.clone()
    segs = (OptSeg * 1)(seg_fn())
    def f(ema):
        segs[0].ema = ema
        _lib.call("drq_adam_pack_step", p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), sc.data_ptr(), p.data_ptr(), tgt.data_ptr(), 0.01, 0.99, segs, 1, s)
    for ema in (0, 1):
        ts = []
        for _ in range(7):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); f(ema); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
        cold = sorted(ts)[3]
        f(ema); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): f(ema)
        e1.record(); torch.cuda.synchronize()
        warm = e0.elapsed_time(e1) * 1e2
        by = n * (bytes_per[1] if ema else bytes_per[0])
        print(f"{name:28s} {'ema ' if ema else 'adam'} n={n:9d} cold {cold:6.1f} us {by / cold * 1e-3:6.0f} GB/s   warm {warm:6.1f} us {by / warm * 1e-3:6.0f} GB/s", flush=True)
n = 4 << 20
run("plain (no bf16 copy)", lambda: OptSeg(OPT_PLAIN, 0, 0, 0, 0, n, None, None), n, (28, 12))
w = TB(1024, 4096, dev, rblk=64)
run("linear 1024 x 4096 (TB copy)", lambda: OptSeg(OPT_LINEAR, 0, 1024, 4096, 0, n, w.ptr(), None), n, (30, 14))
wt = TB(64, 39200, dev, rblk=64)
nt = 50 * 39200
run("trunk 50 x 39200 (TB copy)", lambda: OptSeg(OPT_TRUNK, 0, 50, 0, 0, nt, wt.ptr(), None), nt, (30, 14))
