"""In-kernel clock64 timeline of the bf16 GEMM at the head shapes (debug aid)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from drqv2_b200 import _lib
from drqv2_b200._bf16 import TB, gemm, _strides
from drqv2_b200._lib import TEPI_RELU_BF16, TEPI_F32, TEPI_MASK_BF16, GEMM_KK, GEMM_KMN, GEMM_MNMN
dev = torch.device("cuda")
st = torch.zeros(16, dtype=torch.int64, device=dev)
_lib.call("drq_debug_gemm_stamps", st.data_ptr())
names = ["start", "setup", "-", "issued_all", "land0", "land_last", "acc_done", "epi_done", "exit"]
def run(tag, fn, iters=3):
    for _ in range(iters):
        st.zero_(); fn(); torch.cuda.synchronize()
    v = st.tolist()[:9]
    print(tag, " ".join(f"{n}={v[i]-v[0]}" for i, n in enumerate(names)))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    print("   warm avg %.1f us" % (e0.elapsed_time(e1) * 1e3 / 20))
Bt, H = 256, 1024
x, w, y = TB(Bt, H, dev, batch=4), TB(H, H, dev, batch=4, rblk=64), TB(Bt, H, dev, batch=4)
bias = torch.zeros(4, H, device=dev)
run("fwd 256x1024x1024", lambda: gemm(x.ptr(), x.units, w.ptr(), w.units, GEMM_KK, y.ptr(), y.units, Bt, H, H, TEPI_RELU_BF16, bias=bias.data_ptr()))
run("fwd 4x(256x1024x1024)", lambda: gemm(x.ptr(), x.units, w.ptr(), w.units, GEMM_KK, y.ptr(), y.units, Bt, H, H, TEPI_RELU_BF16, bias=bias.data_ptr(), batch=4, batch_inner=4, strides=_strides((x.stride, w.stride, y.stride, H, 0))))
run("dgrad 2x(256x1024x1024) bn64", lambda: gemm(x.ptr(), x.units, w.ptr(), w.units, GEMM_KMN, y.ptr(), y.units, Bt, H, H, TEPI_MASK_BF16, mask=x.ptr(), units_mask=x.units, batch=2, batch_inner=2, strides=_strides((x.stride, w.stride, y.stride, 0, x.stride))))
run("dgrad 2x(256x1024x1024) bn128", lambda: gemm(x.ptr(), x.units, w.ptr(), w.units, GEMM_KMN, y.ptr(), y.units, Bt, H, H, TEPI_MASK_BF16, mask=x.ptr(), units_mask=x.units, batch=2, batch_inner=2, strides=_strides((x.stride, w.stride, y.stride, 0, x.stride)), bn=128))
dw = torch.zeros(2, H, H, device=dev)
run("wgrad 2x(1024x1024x256) bn128", lambda: gemm(x.ptr(), x.units, y.ptr(), y.units, GEMM_MNMN, dw.data_ptr(), H, H, H, Bt, TEPI_F32, batch=2, batch_inner=2, strides=_strides((x.stride, y.stride, H * H, 0, 0)), bn=128))
run("wgrad 2x(1024x1024x256) bn64", lambda: gemm(x.ptr(), x.units, y.ptr(), y.units, GEMM_MNMN, dw.data_ptr(), H, H, H, Bt, TEPI_F32, batch=2, batch_inner=2, strides=_strides((x.stride, y.stride, H * H, 0, 0)), bn=64))
x56 = TB(Bt, 56, dev); w56 = TB(H, 56, dev, rblk=64)
run("fwd K=56", lambda: gemm(x56.ptr(), x56.units, w56.ptr(), w56.units, GEMM_KK, y.ptr(), y.units, Bt, H, 56, TEPI_RELU_BF16, bias=bias.data_ptr()))
feat = TB(512, 39200, dev); wt = TB(192, 39200, dev, rblk=64)
S = 35
part = torch.zeros(S * 2 * 256 * 128, device=dev)
run("trunk fwd merged (2 x 256 x 128, splitk 35, bn 128)", lambda: gemm(feat.ptr(), feat.units, wt.ptr(row=64), wt.units, GEMM_KK, part.data_ptr(), 128, 256, 128, 39200, TEPI_F32, batch=2, batch_inner=1, bn=128, splitk=S, strides=_strides(outer=(feat.off(row=256), -wt.off(row=64), 256 * 128, 0, 0), split=2 * 256 * 128)))
dz = TB(256, 50, dev)
dwt = torch.zeros(50, 39200, device=dev)
run("trunk wgrad (50 x 39200 x 256)", lambda: gemm(dz.ptr(), dz.units, feat.ptr(), feat.units, GEMM_MNMN, dwt.data_ptr(), 39200, 50, 39200, 256, 3, bn=128))
d4 = torch.zeros(_lib.lib().drq_wb_elems(256), dtype=torch.bfloat16, device=dev)
run("trunk dgrad (256 x 39200 x 50)", lambda: gemm(dz.ptr(), dz.units, wt.ptr(), wt.units, GEMM_KMN, d4.data_ptr(), 256 * 1776 + 128, 256, 39200, 50, 4, mask=feat.ptr(), units_mask=feat.units, bn=128))
