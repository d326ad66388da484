"""Warm CUDA-event timings of the encoder kernels at the B=256 update shapes (10 back-to-back launches)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from drqv2_b200 import _lib
dev = torch.device("cuda"); L = _lib.lib(); s = torch.cuda.current_stream().cuda_stream
Bt = 256
def timeit(fn, n=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n
obs = torch.randint(0, 256, (2 * Bt, 9, 84, 84), dtype=torch.uint8, device=dev)
shift = torch.randint(0, 9, (2 * Bt, 2), dtype=torch.int32, device=dev)
w1 = torch.zeros(L.drq_conv1_w_packed_elems(), dtype=torch.bfloat16, device=dev)
a1 = torch.zeros(L.drq_wb_elems(2 * Bt), dtype=torch.bfloat16, device=dev)
a2 = torch.zeros(L.drq_wb_elems(2 * Bt), dtype=torch.bfloat16, device=dev)
d1 = (torch.randn(L.drq_wb_elems(Bt), device=dev) * 1e-3).to(torch.bfloat16)
d2 = torch.zeros(L.drq_wb_elems(Bt), dtype=torch.bfloat16, device=dev)
ws = torch.zeros(max(L.drq_conv1_wgrad_bf16_ws_floats(), L.drq_conv_wgrad_bf16_ws_floats()), device=dev)
dw1, db = torch.zeros(32, 9, 3, 3, device=dev), torch.zeros(32, device=dev)
dw = torch.zeros(32, 32, 3, 3, device=dev)
wf = torch.zeros(36 * 32 * 8, dtype=torch.bfloat16, device=dev); b1 = torch.zeros(32, device=dev)
print("conv1 fwd N=512   %.1f us" % timeit(lambda: _lib.call("drq_conv1_fwd_bf16", obs.data_ptr(), shift.data_ptr(), w1.data_ptr(), a1.data_ptr(), 2 * Bt, 9, 4, s)))
print("conv1 wgrad N=256 %.1f us" % timeit(lambda: _lib.call("drq_conv1_wgrad_bf16", obs.data_ptr(), shift.data_ptr(), d1.data_ptr(), ws.data_ptr(), dw1.data_ptr(), db.data_ptr(), Bt, 9, 4, s)))
print("conv fwd N=512    %.1f us" % timeit(lambda: _lib.call("drq_conv3x3_fwd_bf16", a1.data_ptr(), wf.data_ptr(), b1.data_ptr(), a2.data_ptr(), 2 * Bt, 39, 0, 0, 0, 0, s)))
print("conv dgrad N=256  %.1f us" % timeit(lambda: _lib.call("drq_conv3x3_dgrad_bf16", d1.data_ptr(), wf.data_ptr(), a1.data_ptr(), 2 * Bt, d2.data_ptr(), Bt, 39, s)))
print("conv wgrad N=256  %.1f us" % timeit(lambda: _lib.call("drq_conv3x3_wgrad_bf16", a1.data_ptr(), 2 * Bt, d1.data_ptr(), ws.data_ptr(), dw.data_ptr(), db.data_ptr(), Bt, 39, s)))
