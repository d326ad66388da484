"""conv3x3 forward (N=512) / data gradient (N=256): the one-pixel-per-row kernel against the four-pixel-column kernel.
CUDA events around single launches, a 256 MB buffer written in front of each (cold L2), and 10 back-to-back (warm);
max |difference| of the outputs."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from drqv2_b200 import _lib
dev = torch.device("cuda"); L = _lib.lib(); s = torch.cuda.current_stream().cuda_stream
Bt = 256
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
a1 = (torch.rand(L.drq_wb_elems(2 * Bt), device=dev) - 0.3).clamp_min(0).to(torch.bfloat16)
d1 = (torch.randn(L.drq_wb_elems(Bt), device=dev) * 1e-3).to(torch.bfloat16)
wf = (torch.randn(36 * 32 * 8, device=dev) * 0.05).to(torch.bfloat16); b1 = torch.randn(32, device=dev) * 0.1
from drqv2_b200._bf16 import TB
for hout in (39, 37, 35):
    outs = {}
    for mode in (0, 2):
        L.drq_set_conv4x1(mode)
        a2 = torch.zeros(L.drq_wb_elems(2 * Bt), dtype=torch.bfloat16, device=dev)
        d2 = torch.zeros(L.drq_wb_elems(Bt), dtype=torch.bfloat16, device=dev)
        f = lambda: _lib.call("drq_conv3x3_fwd_bf16", a1.data_ptr(), wf.data_ptr(), b1.data_ptr(), a2.data_ptr(), 2 * Bt, hout, 0, 0, 0, 0, s)
        g = lambda: _lib.call("drq_conv3x3_dgrad_bf16", d1.data_ptr(), wf.data_ptr(), a1.data_ptr(), 2 * Bt, d2.data_ptr(), Bt, hout, s)
        flops_f = 2 * 2 * Bt * hout * hout * 288 * 32
        tf, tg = cold(f), cold(g)
        print(f"hout {hout} mode {mode}: fwd cold {tf:6.1f} us ({flops_f / tf * 1e-6:6.1f} TFLOP/s) warm {warm(f):6.1f} us | "
              f"dgrad cold {tg:6.1f} us warm {warm(g):6.1f} us", flush=True)
        if hout == 35:
            fb = TB(2 * Bt, hout * hout * 32, dev)
            h = lambda: _lib.call("drq_conv3x3_fwd_bf16", a1.data_ptr(), wf.data_ptr(), b1.data_ptr(), fb.ptr(), 2 * Bt, hout, 2, fb.units, Bt, Bt, s)
            print(f"   TB epilogue: fwd cold {cold(h):6.1f} us warm {warm(h):6.1f} us", flush=True)
        from tests.helpers import nchw_from_wb
        outs[mode] = (nchw_from_wb(a2.view(4, -1, 8), 2 * Bt, hout, hout), nchw_from_wb(d2.view(4, -1, 8), Bt, hout + 2, hout + 2))
    print(f"   max |fwd diff| {(outs[0][0] - outs[2][0]).abs().max().item():.3e} of {outs[0][0].abs().max().item():.3e}; "
          f"max |dgrad diff| {(outs[0][1] - outs[2][1]).abs().max().item():.3e} of {outs[0][1].abs().max().item():.3e}", flush=True)
L.drq_set_conv4x1(2)
st = torch.zeros(16, dtype=torch.int64, device=dev)
L.drq_debug_conv4x1_stamps(st.data_ptr())
a2 = torch.zeros(L.drq_wb_elems(2 * Bt), dtype=torch.bfloat16, device=dev)
d2 = torch.zeros(L.drq_wb_elems(Bt), dtype=torch.bfloat16, device=dev)
for nm, fn in (("fwd", lambda: _lib.call("drq_conv3x3_fwd_bf16", a1.data_ptr(), wf.data_ptr(), b1.data_ptr(), a2.data_ptr(), 2 * Bt, 39, 0, 0, 0, 0, s)),
               ("dgrad", lambda: _lib.call("drq_conv3x3_dgrad_bf16", d1.data_ptr(), wf.data_ptr(), a1.data_ptr(), 2 * Bt, d2.data_ptr(), Bt, 39, s))):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    v = st.tolist()
    print(f"{nm} block 0 clock64 totals: producer wait-stage {v[0]} total {v[3]} | umma wait-acc {v[4]} wait-stage {v[5]} issue {v[6]} total {v[7]} | "
          f"epilogue wait-acc {v[8]} ld {v[9]} rest {v[10]} total {v[11]}")
L.drq_debug_conv4x1_stamps(None)
L.drq_set_conv4x1(1)
