"""conv1 forward (N=512, 9 channels): im2col kernel against the parity-plane kernel, CUDA events, cold L2 and warm."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from drqv2_b200 import _lib
dev = torch.device("cuda"); L = _lib.lib(); s = torch.cuda.current_stream().cuda_stream
N = 512
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def cold(fn, n=7):
    ts = []
    for _ in range(n):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return sorted(ts)[len(ts) // 2]
def warm(fn, n=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n
obs = torch.randint(0, 256, (N, 9, 84, 84), dtype=torch.uint8, device=dev)
shift = torch.randint(0, 9, (N, 2), dtype=torch.int32, device=dev)
w = (torch.rand(32, 9, 3, 3, device=dev) - 0.5) * 0.3
b = (torch.rand(32, device=dev) - 0.5) * 0.1
wp = torch.zeros(L.drq_conv1_w_packed_elems(), dtype=torch.bfloat16, device=dev)
_lib.call("drq_pack_conv1_w_bf16", w.data_ptr(), b.data_ptr(), wp.data_ptr(), 9, s)
outs = {}
for mode in (0, 2):
    L.drq_set_conv1_planes(mode)
    out = torch.zeros(L.drq_wb_elems(N), dtype=torch.bfloat16, device=dev)
    f = lambda: _lib.call("drq_conv1_fwd_bf16", obs.data_ptr(), shift.data_ptr(), wp.data_ptr(), out.data_ptr(), N, 9, 4, s)
    print(f"conv1 fwd mode {mode}: cold {cold(f):6.1f} us  warm {warm(f):6.1f} us", flush=True)
    outs[mode] = out.float()
print("max |diff|", (outs[0] - outs[2]).abs().max().item(), "of", outs[0].abs().max().item())
L.drq_set_conv1_planes(1)
st = torch.zeros(16, dtype=torch.int64, device=dev)
L.drq_debug_conv1_stamps(st.data_ptr())
out = torch.zeros(L.drq_wb_elems(N), dtype=torch.bfloat16, device=dev)
for _ in range(3):
    _lib.call("drq_conv1_fwd_bf16", obs.data_ptr(), shift.data_ptr(), wp.data_ptr(), out.data_ptr(), N, 9, 4, s)
torch.cuda.synchronize()
v = st.tolist()
print(f"planes kernel, block 0, builder group 0 thread 0 (clock64 totals): wait rows {v[0]} re-pitch+barrier {v[1]} wait stage {v[2]} build {v[3]} "
      f"fence+arrive {v[4]} total {v[5]} | umma wait-acc {v[6]} wait-stage {v[7]} total {v[8]} | producer wait-free-stage {v[9]} total {v[10]}")
L.drq_debug_conv1_stamps(None)
