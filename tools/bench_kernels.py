"""Time individual kernels at the bench sizes (B=256) with CUDA events."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from drqv2_b200 import _lib  # noqa: E402

dev = torch.device("cuda")
s = torch.cuda.current_stream().cuda_stream
L = _lib.lib()


def timeit(fn, iters=20):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3   # us


B = 256
NB = 2 * B
MAC = {39: 14_017_536, 37: 12_616_704, 35: 11_289_600}
x = (torch.rand(L.drq_wb_elems(NB), device=dev) - 0.3).clamp_min(0).to(torch.bfloat16)
y = torch.zeros(L.drq_wb_elems(NB), dtype=torch.bfloat16, device=dev)
d = ((torch.rand(L.drq_wb_elems(B), device=dev) - 0.5) * 1e-3).to(torch.bfloat16)
d2 = torch.zeros(L.drq_wb_elems(B), dtype=torch.bfloat16, device=dev)
w = (torch.rand(32, 32, 3, 3, device=dev) - 0.5) * 0.1
bias = torch.zeros(32, device=dev)
wf = torch.zeros(36 * 32 * 8, dtype=torch.bfloat16, device=dev)
wd = torch.zeros_like(wf)
_lib.call("drq_pack_conv_w_bf16", w.data_ptr(), wf.data_ptr(), wd.data_ptr(), s)
ws = torch.zeros(L.drq_conv_wgrad_bf16_ws_floats(), device=dev)
dw = torch.zeros(32, 32, 3, 3, device=dev)
db = torch.zeros(32, device=dev)
for hout in (39, 37, 35):
    t = timeit(lambda: _lib.call("drq_conv3x3_fwd_bf16", x.data_ptr(), wf.data_ptr(), bias.data_ptr(), y.data_ptr(), NB, hout, 0, s))
    print(f"fwd   hout={hout} N={NB}: {t:8.1f} us  {2 * MAC[hout] * NB / t / 1e6:8.1f} TFLOP/s")
    t = timeit(lambda: _lib.call("drq_conv3x3_dgrad_bf16", d.data_ptr(), wd.data_ptr(), x.data_ptr(), NB, d2.data_ptr(), B, hout, s))
    print(f"dgrad hout={hout} N={B}: {t:8.1f} us  {2 * MAC[hout] * B / t / 1e6:8.1f} TFLOP/s")
    t = timeit(lambda: _lib.call("drq_conv3x3_wgrad_bf16", x.data_ptr(), NB, d.data_ptr(), ws.data_ptr(), dw.data_ptr(), db.data_ptr(), B, hout, s))
    print(f"wgrad hout={hout} N={B}: {t:8.1f} us  {2 * MAC[hout] * B / t / 1e6:8.1f} TFLOP/s")
