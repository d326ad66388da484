"""Runs conv3x3 forward (N=512), dgrad and wgrad (N=256) at the layer-2 shape (driver for ncu)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from drqv2_b200 import _lib
dev = torch.device("cuda"); L = _lib.lib(); s = torch.cuda.current_stream().cuda_stream
Bt = 256
a1 = (torch.rand(L.drq_wb_elems(2 * Bt), device=dev) - 0.3).clamp_min(0).to(torch.bfloat16)
a2 = torch.zeros(L.drq_wb_elems(2 * Bt), dtype=torch.bfloat16, device=dev)
d1 = (torch.randn(L.drq_wb_elems(Bt), device=dev) * 1e-3).to(torch.bfloat16)
d2 = torch.zeros(L.drq_wb_elems(Bt), dtype=torch.bfloat16, device=dev)
ws = torch.zeros(L.drq_conv_wgrad_bf16_ws_floats(), device=dev)
dw, db = torch.zeros(32, 32, 3, 3, device=dev), torch.zeros(32, device=dev)
wf = (torch.randn(36 * 32 * 8, device=dev) * 0.05).to(torch.bfloat16); b1 = torch.zeros(32, device=dev)
for _ in range(3):
    _lib.call("drq_conv3x3_fwd_bf16", a1.data_ptr(), wf.data_ptr(), b1.data_ptr(), a2.data_ptr(), 2 * Bt, 39, 0, 0, 0, 0, s)
    _lib.call("drq_conv3x3_dgrad_bf16", d1.data_ptr(), wf.data_ptr(), a1.data_ptr(), 2 * Bt, d2.data_ptr(), Bt, 39, s)
    _lib.call("drq_conv3x3_wgrad_bf16", a1.data_ptr(), 2 * Bt, d1.data_ptr(), ws.data_ptr(), dw.data_ptr(), db.data_ptr(), Bt, 39, s)
torch.cuda.synchronize()
print("ok")
