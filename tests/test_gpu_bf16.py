"""GPU parity tests of the bf16 tensor-core (tcgen05) kernels against fp32/fp64 torch math on
bf16-rounded operands.  Tolerances: the products are exact in fp32, accumulation is fp32 in
TMEM, so the only rounding beyond the reference is the final bf16 store (2^-9 relative)."""
import numpy as np
import pytest
import torch

from tests.helpers import GUARD, PLB, nchw_from_wb, wb_from_nchw

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda")


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _bf(x):
    return x.to(torch.bfloat16).float()


def _pack_w(w, dev):
    from drqv2_b200 import _lib
    wf = torch.zeros(36 * 32 * 8, dtype=torch.bfloat16, device=dev)
    wd = torch.zeros_like(wf)
    _lib.call("drq_pack_conv_w_bf16", w.data_ptr(), wf.data_ptr(), wd.data_ptr(), _stream())
    return wf, wd


@pytest.mark.parametrize("hout,N", [(39, 3), (37, 5), (35, 2)])
def test_conv3x3_fwd_bf16(dev, hout, N):
    from drqv2_b200 import _lib
    g = torch.Generator().manual_seed(hout)
    hin = hout + 2
    x = torch.rand(N, 32, hin, hin, generator=g).to(dev)
    w = ((torch.rand(32, 32, 3, 3, generator=g) * 2 - 1) * 0.1).to(dev)
    b = ((torch.rand(32, generator=g) * 2 - 1) * 0.1).to(dev)
    xin = wb_from_nchw(x)
    wf, _ = _pack_w(w, dev)
    out = torch.zeros(_lib.lib().drq_wb_elems(N), dtype=torch.bfloat16, device=dev)
    _lib.call("drq_conv3x3_fwd_bf16", xin.data_ptr(), wf.data_ptr(), b.data_ptr(), out.data_ptr(), N, hout, 0, _stream())
    torch.cuda.synchronize()
    want = torch.relu(torch.nn.functional.conv2d(_bf(x).double(), _bf(w).double(), b.double()))
    got = nchw_from_wb(out.view(4, -1, 8), N, hout, hout).double()
    err = (got - want).abs().max().item()
    assert err <= 2 ** -8 * want.abs().max().item() + 1e-6, err
    # compact NHWC feature output
    feat = torch.zeros(N, hout * hout, 32, dtype=torch.bfloat16, device=dev)
    _lib.call("drq_conv3x3_fwd_bf16", xin.data_ptr(), wf.data_ptr(), b.data_ptr(), feat.data_ptr(), N, hout, 1, _stream())
    torch.cuda.synchronize()
    got2 = feat.float().view(N, hout, hout, 32).permute(0, 3, 1, 2).double()
    assert torch.equal(got2, got)


@pytest.mark.parametrize("hout,N", [(39, 2), (35, 3)])
def test_conv3x3_dgrad_bf16(dev, hout, N):
    from drqv2_b200 import _lib
    g = torch.Generator().manual_seed(100 + hout)
    hin = hout + 2
    dout = ((torch.rand(N, 32, hout, hout, generator=g) * 2 - 1) * 1e-3).to(dev)
    act = (torch.rand(N + 1, 32, hin, hin, generator=g) - 0.4).clamp_min(0).to(dev)   # post-ReLU input act, extra image
    w = ((torch.rand(32, 32, 3, 3, generator=g) * 2 - 1) * 0.1).to(dev)
    d_wb = wb_from_nchw(dout)
    a_wb = wb_from_nchw(act)
    _, wd = _pack_w(w, dev)
    din = torch.full((_lib.lib().drq_wb_elems(N),), 7.0, dtype=torch.bfloat16, device=dev)
    _lib.call("drq_conv3x3_dgrad_bf16", d_wb.data_ptr(), wd.data_ptr(), a_wb.data_ptr(), N + 1, din.data_ptr(), N, hout, _stream())
    torch.cuda.synchronize()
    want = torch.nn.functional.conv_transpose2d(_bf(dout).double(), _bf(w).double())
    want = want * (_bf(act[:N]).double() > 0)
    got = nchw_from_wb(din.view(4, -1, 8), N, hin, hin).double()
    err = (got - want).abs().max().item()
    assert err <= 2 ** -8 * want.abs().max().item() + 1e-9, err
    # columns beyond the valid width are written as exact zeros (consumed by wgrad/dgrad below)
    full = nchw_from_wb(din.view(4, -1, 8), N, 41, 41)
    assert torch.count_nonzero(full[:, :, :hin, hin:]) == 0


@pytest.mark.parametrize("hout,N", [(39, 3), (35, 7)])
def test_conv3x3_wgrad_bf16(dev, hout, N):
    from drqv2_b200 import _lib
    g = torch.Generator().manual_seed(200 + hout)
    hin = hout + 2
    x = torch.rand(N + 2, 32, hin, hin, generator=g).to(dev)                   # input activation (more images than d)
    d = ((torch.rand(N, 32, hout, hout, generator=g) * 2 - 1) * 1e-2).to(dev)
    x_wb, d_wb = wb_from_nchw(x), wb_from_nchw(d)
    ws = torch.zeros(_lib.lib().drq_conv_wgrad_bf16_ws_floats(), device=dev)
    dw = torch.zeros(32, 32, 3, 3, device=dev)
    db = torch.zeros(32, device=dev)
    _lib.call("drq_conv3x3_wgrad_bf16", x_wb.data_ptr(), N + 2, d_wb.data_ptr(), ws.data_ptr(), dw.data_ptr(),
              db.data_ptr(), N, hout, _stream())
    torch.cuda.synchronize()
    xr, dr = _bf(x[:N]).double(), _bf(d).double()
    want_w = torch.nn.grad.conv2d_weight(xr, (32, 32, 3, 3), dr)
    want_b = dr.sum(dim=(0, 2, 3))
    assert (dw.double() - want_w).abs().max().item() <= 1e-5 * want_w.abs().max().item() + 1e-7
    assert (db.double() - want_b).abs().max().item() <= 1e-5 * want_b.abs().max().item() + 1e-7
