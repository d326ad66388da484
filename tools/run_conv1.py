"""Runs conv1 forward (N=512) and wgrad (N=256) a few times (driver for ncu)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from drqv2_b200 import _lib
dev = torch.device("cuda"); L = _lib.lib(); s = torch.cuda.current_stream().cuda_stream
Bt = 256
obs = torch.randint(0, 256, (2 * Bt, 9, 84, 84), dtype=torch.uint8, device=dev)
shift = torch.randint(0, 9, (2 * Bt, 2), dtype=torch.int32, device=dev)
w = torch.randn(32, 9, 3, 3, device=dev) * 0.1; b = torch.zeros(32, device=dev)
w1 = torch.zeros(L.drq_conv1_w_packed_elems(), dtype=torch.bfloat16, device=dev)
_lib.call("drq_pack_conv1_w_bf16", w.data_ptr(), b.data_ptr(), w1.data_ptr(), 9, s)
a1 = torch.zeros(L.drq_wb_elems(2 * Bt), dtype=torch.bfloat16, device=dev)
d1 = (torch.randn(L.drq_wb_elems(Bt), device=dev) * 1e-3).to(torch.bfloat16)
ws = torch.zeros(L.drq_conv1_wgrad_bf16_ws_floats(), device=dev)
dw, db = torch.zeros(32, 9, 3, 3, device=dev), torch.zeros(32, device=dev)
for _ in range(3):
    _lib.call("drq_conv1_fwd_bf16", obs.data_ptr(), shift.data_ptr(), w1.data_ptr(), a1.data_ptr(), 2 * Bt, 9, 4, s)
    _lib.call("drq_conv1_wgrad_bf16", obs.data_ptr(), shift.data_ptr(), d1.data_ptr(), ws.data_ptr(), dw.data_ptr(), db.data_ptr(), Bt, 9, 4, s)
torch.cuda.synchronize()
print("ok")
