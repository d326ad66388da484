"""Per-role wait/work cycle totals of the tensor-core conv kernels (block 0) at B=256 shapes."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from drqv2_b200 import _lib
dev = torch.device("cuda")
L = _lib.lib()
s = torch.cuda.current_stream().cuda_stream
st = torch.zeros(16, dtype=torch.int64, device=dev)
_lib.call("drq_debug_conv_stamps", st.data_ptr())
Bt = 256
zb = lambda n: (torch.rand(n, device=dev) - 0.3).clamp_min(0).to(torch.bfloat16)
xw = zb(L.drq_wb_elems(2 * Bt)); yw = torch.zeros(L.drq_wb_elems(2 * Bt), dtype=torch.bfloat16, device=dev)
wf = torch.zeros(36 * 32 * 8, dtype=torch.bfloat16, device=dev); b1 = torch.zeros(32, device=dev)
dd = zb(L.drq_wb_elems(Bt)); d2 = torch.zeros(L.drq_wb_elems(Bt), dtype=torch.bfloat16, device=dev)
names = ["prod wait empty", "prod total", "mma wait tempty", "mma wait full", "mma total", "epi wait tfull"]
def run(tag, fn):
    for _ in range(3):
        st.zero_(); fn(); torch.cuda.synchronize()
    v = st.tolist()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    _lib.call("drq_debug_conv_stamps", None)
    fn(); e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    _lib.call("drq_debug_conv_stamps", st.data_ptr())
    print(tag, "| %.1f us |" % (e0.elapsed_time(e1) * 100), ", ".join(f"{n}={v[i]}" for i, n in enumerate(names)))
run("fwd N=512 h39", lambda: _lib.call("drq_conv3x3_fwd_bf16", xw.data_ptr(), wf.data_ptr(), b1.data_ptr(), yw.data_ptr(), 2 * Bt, 39, 0, 0, 0, 0, s))
run("dgrad N=256 h39", lambda: _lib.call("drq_conv3x3_dgrad_bf16", dd.data_ptr(), wf.data_ptr(), xw.data_ptr(), 2 * Bt, d2.data_ptr(), Bt, 39, s))

# conv1 forward builders
obs = torch.randint(0, 256, (2 * Bt, 9, 84, 84), dtype=torch.uint8, device=dev)
shift = torch.randint(0, 9, (2 * Bt, 2), dtype=torch.int32, device=dev)
w1 = torch.zeros(L.drq_conv1_w_packed_elems(), dtype=torch.bfloat16, device=dev)
a1 = torch.zeros(L.drq_wb_elems(2 * Bt), dtype=torch.bfloat16, device=dev)
_lib.call("drq_debug_conv_stamps", None)
_lib.call("drq_debug_conv1_stamps", st.data_ptr())
f1 = lambda: _lib.call("drq_conv1_fwd_bf16", obs.data_ptr(), shift.data_ptr(), w1.data_ptr(), a1.data_ptr(), 2 * Bt, 9, 4, s)
for _ in range(3):
    st.zero_(); f1(); torch.cuda.synchronize()
v = st.tolist()
print("conv1 fwd N=512 builder thread 0  : rows-wait %d pad %d bar %d tile-wait %d build %d" % tuple(v[:5]))
print("conv1 fwd N=512 builder thread 128: rows-wait %d pad %d bar %d tile-wait %d build %d" % tuple(v[8:13]))
_lib.call("drq_debug_conv1_stamps", None)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
f1(); e0.record()
for _ in range(10): f1()
e1.record(); torch.cuda.synchronize()
print("conv1 fwd %.1f us" % (e0.elapsed_time(e1) * 100))
