"""Smallest launches of every tcgen05 pipeline kernel (conv1 fwd / wgrad, conv3x3 fwd / dgrad / wgrad, the three GEMM
operand modes, the policy head) for `compute-sanitizer --tool racecheck|memcheck`.  Several tiles per CTA
where it is cheap, so that the shared-memory stage rings are re-used."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from drqv2_b200 import _lib  # noqa: E402
from drqv2_b200._bf16 import TB, gemm, _strides  # noqa: E402
from drqv2_b200._lib import GEMM_KK, GEMM_KMN, GEMM_MNMN, TEPI_F32, TEPI_MASK_BF16, TEPI_RELU_BF16  # noqa: E402

dev = torch.device("cuda")
L = _lib.lib()
s = torch.cuda.current_stream().cuda_stream
N = int(sys.argv[1]) if len(sys.argv) > 1 else 24          # 24 images x 13 tiles = 312 tiles: > 296 CTAs
g = torch.Generator(device="cuda").manual_seed(0)
rnd = lambda n, sc=1.0: (torch.rand(n, device=dev, generator=g) - 0.3).clamp_min(0).mul(sc).to(torch.bfloat16)
a1, a2 = rnd(L.drq_wb_elems(N)), torch.zeros(L.drq_wb_elems(N), dtype=torch.bfloat16, device=dev)
d1, d2 = rnd(L.drq_wb_elems(N), 1e-3), torch.zeros(L.drq_wb_elems(N), dtype=torch.bfloat16, device=dev)
wf = rnd(36 * 32 * 8, 0.05)
b1 = torch.zeros(32, device=dev)
ws = torch.zeros(L.drq_conv_wgrad_bf16_ws_floats(), device=dev)
dw, db = torch.zeros(32, 32, 3, 3, device=dev), torch.zeros(32, device=dev)
_lib.call("drq_conv3x3_fwd_bf16", a1.data_ptr(), wf.data_ptr(), b1.data_ptr(), a2.data_ptr(), N, 39, 0, 0, 0, 0, s)
_lib.call("drq_conv3x3_dgrad_bf16", d1.data_ptr(), wf.data_ptr(), a1.data_ptr(), N, d2.data_ptr(), N, 39, s)
_lib.call("drq_conv3x3_wgrad_bf16", a1.data_ptr(), N, d1.data_ptr(), ws.data_ptr(), dw.data_ptr(), db.data_ptr(), N, 39, s)
obs = torch.randint(0, 256, (N, 9, 84, 84), dtype=torch.uint8, device=dev, generator=g)
shift = torch.randint(0, 9, (N, 2), dtype=torch.int32, device=dev, generator=g)
w = torch.randn(32, 9, 3, 3, device=dev, generator=g) * 0.1
w1 = torch.zeros(L.drq_conv1_w_packed_elems(), dtype=torch.bfloat16, device=dev)
_lib.call("drq_pack_conv1_w_bf16", w.data_ptr(), b1.data_ptr(), w1.data_ptr(), 9, s)
ws1 = torch.zeros(L.drq_conv1_wgrad_bf16_ws_floats(), device=dev)
dw1 = torch.zeros(32, 9, 3, 3, device=dev)
_lib.call("drq_conv1_fwd_bf16", obs.data_ptr(), shift.data_ptr(), w1.data_ptr(), a2.data_ptr(), N, 9, 4, s)
_lib.call("drq_conv1_wgrad_bf16", obs.data_ptr(), shift.data_ptr(), d1.data_ptr(), ws1.data_ptr(), dw1.data_ptr(), db.data_ptr(), N, 9, 4, s)
# GEMMs: K-major forward (K = 1024: the stage ring wraps), K/MN data gradient with mask, MN/MN weight gradient
M, H, I = 256, 1024, 56
x, wt, y = TB(M, H, dev), TB(H, H, dev, rblk=64), TB(M, H, dev)
x.buf.copy_(rnd(x.buf.numel())); wt.buf.copy_(rnd(wt.buf.numel(), 0.05))
bias = torch.zeros(H, device=dev)
gemm(x.ptr(), x.units, wt.ptr(), wt.units, GEMM_KK, y.ptr(), y.units, M, H, H, TEPI_RELU_BF16, bias=bias.data_ptr())
dx = TB(M, H, dev)
gemm(y.ptr(), y.units, wt.ptr(), wt.units, GEMM_KMN, dx.ptr(), dx.units, M, H, H, TEPI_MASK_BF16, mask=x.ptr(), units_mask=x.units)
dW = torch.zeros(H, H, device=dev)
gemm(y.ptr(), y.units, x.ptr(), x.units, GEMM_MNMN, dW.data_ptr(), H, H, H, M, TEPI_F32, bn=128)
part = torch.zeros(8, M, I, device=dev)
w0 = TB(H, I, dev, rblk=64); w0.buf.copy_(rnd(w0.buf.numel(), 0.05))
gemm(y.ptr(), y.units, w0.ptr(), w0.units, GEMM_KMN, part.data_ptr(), I, M, I, H, TEPI_F32, splitk=8, strides=_strides(split=M * I))
torch.cuda.synchronize()
print("sanitize_small ok: dW sum", float(dW.abs().sum()), "conv sums", float(a2.float().abs().sum()), float(dw.abs().sum()), float(dw1.abs().sum()))
