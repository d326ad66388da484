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


@pytest.fixture(params=[0, 2], ids=["pixel_rows", "column_rows"])
def conv_mode(request):
    """both conv3x3 forward / data-gradient kernels at every size: 0 = one pixel per accumulator row (conv_tc.cu),
    2 = a column of four pixels per row (conv4x1_tc.cu; the default mode 1 picks it from 48 images up)"""
    from drqv2_b200 import _lib
    prev = _lib.lib().drq_set_conv4x1(request.param)
    yield request.param
    _lib.lib().drq_set_conv4x1(prev)


@pytest.mark.parametrize("hout,N", [(39, 3), (37, 5), (35, 2)])
def test_conv3x3_fwd_bf16(dev, hout, N, conv_mode):
    from drqv2_b200 import _lib
    g = torch.Generator().manual_seed(hout)
    hin = hout + 2
    x = torch.rand(N, 32, hin, hin, generator=g).to(dev)
    w = ((torch.rand(32, 32, 3, 3, generator=g) * 2 - 1) * 0.1).to(dev)
    b = ((torch.rand(32, generator=g) * 2 - 1) * 0.1).to(dev)
    xin = wb_from_nchw(x)
    wf, _ = _pack_w(w, dev)
    out = torch.zeros(_lib.lib().drq_wb_elems(N), dtype=torch.bfloat16, device=dev)
    _lib.call("drq_conv3x3_fwd_bf16", xin.data_ptr(), wf.data_ptr(), b.data_ptr(), out.data_ptr(), N, hout, 0, 0, 0, 0, _stream())
    torch.cuda.synchronize()
    want = torch.relu(torch.nn.functional.conv2d(_bf(x).double(), _bf(w).double(), b.double()))
    got = nchw_from_wb(out.view(4, -1, 8), N, hout, hout).double()
    err = (got - want).abs().max().item()
    assert err <= 2 ** -8 * want.abs().max().item() + 1e-6, err
    # compact NHWC feature output
    feat = torch.zeros(N, hout * hout, 32, dtype=torch.bfloat16, device=dev)
    _lib.call("drq_conv3x3_fwd_bf16", xin.data_ptr(), wf.data_ptr(), b.data_ptr(), feat.data_ptr(), N, hout, 1, 0, 0, 0, _stream())
    torch.cuda.synchronize()
    got2 = feat.float().view(N, hout, hout, 32).permute(0, 3, 1, 2).double()
    assert torch.equal(got2, got)
    # TB feature matrix (rows = images, channel-group-major feature order); the second "half" starts on its own row block
    from drqv2_b200._bf16 import TB
    half = N // 2
    fb = TB(128 + N, hout * hout * 32, dev)
    _lib.call("drq_conv3x3_fwd_bf16", xin.data_ptr(), wf.data_ptr(), b.data_ptr(), fb.ptr(), N, hout, 2, fb.units,
              half, 128, _stream())
    torch.cuda.synchronize()
    rows = torch.cat([fb.view()[:half], fb.view()[128:128 + N - half]]).float()
    # feature order of the bf16 encoder output: (c/8)*hw*8 + yx*8 + c%8
    got3 = rows[:, :hout * hout * 32].view(N, 4, hout * hout, 8).permute(0, 1, 3, 2).reshape(N, 32, hout, hout).double()
    assert torch.equal(got3, got)


@pytest.mark.parametrize("hout,N", [(39, 2), (35, 3)])
def test_conv3x3_dgrad_bf16(dev, hout, N, conv_mode):
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


def _tb(x, rblk=128):
    """[rows][feats] (or [batch][rows][feats]) fp32 -> TB bf16 buffer object (drqv2_b200._bf16.TB)."""
    from drqv2_b200._bf16 import TB
    x3 = x if x.dim() == 3 else x.unsqueeze(0)
    tb = TB(x3.shape[1], x3.shape[2], x.device, batch=x3.shape[0], rblk=rblk)
    for z in range(x3.shape[0]):
        tb.load(x3[z], z)
    return tb


def _gemm_bf16(A, B, mode, C, ldc, M, N, K, epi, bias=None, mask=None, acc=0, batch=1, bs_c=0, split_stride=0,
               splitk=1, bn=64, n_store=0, a_row0=0):
    """A, B, mask: TB objects; C: TB object (bf16 epilogues) or fp32 tensor."""
    from drqv2_b200 import _lib
    from drqv2_b200._bf16 import TB, _strides
    c_ptr = C.ptr() if isinstance(C, TB) else C.data_ptr()
    st = _strides((A.stride if A.batch > 1 else 0, B.stride if B.batch > 1 else 0,
                   C.stride if isinstance(C, TB) else bs_c, 0,
                   0 if mask is None else (mask.stride if mask.batch > 1 else 0)), split=split_stride)
    _lib.call("drq_gemm_bf16", A.ptr(row=a_row0), A.units, B.ptr(), B.units, mode, c_ptr, ldc, n_store,
              None if bias is None else bias.data_ptr(), None if mask is None else mask.ptr(),
              0 if mask is None else mask.units, M, N, K, epi, acc, batch, batch, st, splitk, bn, _stream())


KK, KMN, MNMN = 0, 1, 2


@pytest.mark.parametrize("M,N,K", [(256, 1024, 56), (37, 130, 1024), (256, 50, 1000), (300, 256, 256)])
def test_gemm_bf16_kmajor_relu_and_f32(dev, M, N, K):
    """Linear forward: y = relu(x W^T + b) (TB bf16 out) and plain fp32 out."""
    from drqv2_b200._bf16 import TB
    g = torch.Generator().manual_seed(M + N + K)
    x = (torch.rand(M, K, generator=g) - 0.5).to(dev)
    w = ((torch.rand(N, K, generator=g) - 0.5) * 0.2).to(dev)
    b = (torch.rand(N, generator=g) - 0.5).to(dev)
    xb, wb = _tb(x), _tb(w, 64)
    want = _bf(x).double() @ _bf(w).double().T + b.double()
    y = TB(M, N, dev)
    y.buf.fill_(7.0)
    _gemm_bf16(xb, wb, KK, y, y.units, M, N, K, 1, bias=b, n_store=y.units * 8)
    yf = torch.zeros(M, N, device=dev)
    _gemm_bf16(xb, wb, KK, yf, N, M, N, K, 0, bias=b)
    torch.cuda.synchronize()
    scale = want.abs().max().item()
    assert (yf.double() - want).abs().max().item() <= 2e-5 * scale
    assert (y.dense().double() - torch.relu(want)).abs().max().item() <= 2 ** -8 * scale
    # feature padding columns [N, ceil16(N)) are written as zeros
    assert torch.count_nonzero(y.view()[:M, N:]) == 0
    # accumulate into C
    _gemm_bf16(xb, wb, KK, yf, N, M, N, K, 0, acc=1)
    torch.cuda.synchronize()
    assert (yf.double() - (2 * want - b.double())).abs().max().item() <= 4e-5 * scale


def test_gemm_bf16_dgrad_wgrad_layouts(dev):
    """dgrad: dx = (dy W) * (x_act > 0) with the weight contracted over its rows; wgrad: dW = dy^T x
    with both activations contracted over their rows; twin-head batching; split-K partials; A rows at
    a row-block offset."""
    from drqv2_b200._bf16 import TB
    g = torch.Generator().manual_seed(5)
    Bt, H, I = 256, 1024, 56
    dy = ((torch.rand(2, Bt, H, generator=g) - 0.5) * 1e-2).to(dev)
    w = ((torch.rand(2, H, I, generator=g) - 0.5) * 0.2).to(dev)
    xact = (torch.rand(2, Bt, I, generator=g) - 0.5).clamp_min(0).to(dev)
    dyb, wb, xb = _tb(dy), _tb(w, 64), _tb(xact)
    # dgrad, batch of 2 heads
    dx = TB(Bt, I, dev, batch=2)
    _gemm_bf16(dyb, wb, KMN, dx, dx.units, Bt, I, H, 2, mask=xb, batch=2)
    torch.cuda.synchronize()
    want = (_bf(dy).double() @ _bf(w).double()) * (xact.double() > 0)
    got = torch.stack([dx.dense(0), dx.dense(1)]).double()
    assert (got - want).abs().max().item() <= 2 ** -8 * want.abs().max().item()
    # the 128-wide N tile (hidden -> hidden data gradient)
    w2 = ((torch.rand(H, 384, generator=g) - 0.5) * 0.2).to(dev)
    m2 = (torch.rand(Bt, 384, generator=g) - 0.5).clamp_min(0).to(dev)
    dx2 = TB(Bt, 384, dev)
    dy0 = TB(Bt, H, dev)
    dy0.load(dy[0])
    _gemm_bf16(dy0, _tb(w2, 64), KMN, dx2, dx2.units, Bt, 384, H, 2, mask=_tb(m2), bn=128)
    torch.cuda.synchronize()
    want2 = (_bf(dy[0]).double() @ _bf(w2).double()) * (m2.double() > 0)
    assert (dx2.dense().double() - want2).abs().max().item() <= 2 ** -8 * want2.abs().max().item()
    # wgrad: dW[h][i] = sum_b dy[b][h] x[b][i]
    dw = torch.zeros(2, H, I, device=dev)
    _gemm_bf16(dyb, xb, MNMN, dw, I, H, I, Bt, 0, batch=2, bs_c=H * I)
    torch.cuda.synchronize()
    want_w = _bf(dy).double().transpose(1, 2) @ _bf(xact).double()
    assert (dw.double() - want_w).abs().max().item() <= 2e-5 * want_w.abs().max().item()
    # wgrad with the 128-wide N tile and K not a multiple of the row block
    Bs = 200
    dys, xs = dy[0, :Bs, :300].contiguous(), dy[1, :Bs, :260].contiguous()
    dw2 = torch.zeros(300, 260, device=dev)
    _gemm_bf16(_tb(dys), _tb(xs), MNMN, dw2, 260, 300, 260, Bs, 0, bn=128)
    torch.cuda.synchronize()
    want_w2 = _bf(dys).double().T @ _bf(xs).double()
    assert (dw2.double() - want_w2).abs().max().item() <= 2e-5 * want_w2.abs().max().item()
    # split-K partials sum to the full product; A rows taken at a row-block offset (the "next" half)
    K = 39200
    feat = (torch.rand(256, K, generator=g) - 0.3).clamp_min(0).to(dev)
    wt = ((torch.rand(50, K, generator=g) - 0.5) * 0.01).to(dev)
    fb, wtb = _tb(feat), _tb(wt, 64)
    S = 35
    part = torch.zeros(S, 100, 50, device=dev)
    _gemm_bf16(fb, wtb, KK, part, 50, 100, 50, K, 0, splitk=S, split_stride=100 * 50, a_row0=128)
    torch.cuda.synchronize()
    want_t = _bf(feat[128:228]).double() @ _bf(wt).double().T
    assert (part.double().sum(0) - want_t).abs().max().item() <= 2e-5 * want_t.abs().max().item()


def test_trunk_weight_pack_and_epilogues(dev):
    """Encoder-order trunk weight pack; wgrad epilogue writes the reference [F][39200] order;
    dgrad epilogue masks by the feature and scatters into conv4's WB gradient plane."""
    from drqv2_b200 import _lib
    from drqv2_b200._bf16 import TB
    from drqv2_b200._lib import PLB as PLB_
    g = torch.Generator().manual_seed(9)
    Fd, Bt, K = 50, 6, 39200
    w = ((torch.rand(Fd, K, generator=g) - 0.5) * 0.02).to(dev)
    wp = TB(Fd, K, dev, rblk=64)
    _lib.call("drq_pack_trunk_tb", w.data_ptr(), wp.ptr(), Fd, _stream())
    # reference order -> encoder-output order: column (c/8)*9800 + (y*35+x)*8 + c%8 holds w[:, c*1225 + y*35 + x]
    w_nhwc = w.view(Fd, 4, 8, 1225).permute(0, 1, 3, 2).reshape(Fd, K)
    assert torch.equal(wp.dense(), _bf(w_nhwc))
    # plain Linear weight pack (more than one 64-row block)
    w2 = ((torch.rand(70, 50, generator=g) - 0.5)).to(dev)
    w2p = TB(70, 50, dev, rblk=64)
    _lib.call("drq_pack_linear_tb", w2.data_ptr(), w2p.ptr(), 70, 50, _stream())
    assert torch.equal(w2p.dense(), _bf(w2))
    feat_nchw = (torch.rand(Bt, 32, 35, 35, generator=g) - 0.4).clamp_min(0).to(dev)
    feat = _tb(feat_nchw.view(Bt, 4, 8, 1225).permute(0, 1, 3, 2).reshape(Bt, K))     # encoder-output feature order
    dz = ((torch.rand(Bt, Fd, generator=g) - 0.5) * 1e-2).to(dev)
    dzb = _tb(dz)
    # wgrad: dW[f][ref(n)] = sum_b dz[b][f] feat[b][n]
    dw = torch.zeros(Fd, K, device=dev)
    _gemm_bf16(dzb, feat, MNMN, dw, K, Fd, K, Bt, 3, bn=128)
    torch.cuda.synchronize()
    want_w = _bf(dz).double().T @ _bf(feat_nchw).double().reshape(Bt, K)
    assert (dw.double() - want_w).abs().max().item() <= 2e-5 * want_w.abs().max().item()
    # dgrad: d4pre = (dz W) * (feat > 0) scattered to WB
    d4 = torch.zeros(_lib.lib().drq_wb_elems(Bt), dtype=torch.bfloat16, device=dev)
    cs = Bt * PLB_ + 128
    _gemm_bf16(dzb, wp, KMN, d4, cs, Bt, K, Fd, 4, mask=feat, bn=128)
    torch.cuda.synchronize()
    want_d = (_bf(dz).double() @ _bf(w).double()).view(Bt, 32, 35, 35) * (feat_nchw.double() > 0)
    got = nchw_from_wb(d4.view(4, -1, 8), Bt, 35, 35).double()
    assert (got - want_d).abs().max().item() <= 2 ** -8 * want_d.abs().max().item()


@pytest.fixture(params=[0, 2], ids=["im2col", "planes"])
def conv1_mode(request):
    """both conv1 forward kernels at every size: 0 = im2col tile per output position, 2 = parity planes + tap offsets
    (the default mode 1 picks it from 48 images up)"""
    from drqv2_b200 import _lib
    prev = _lib.lib().drq_set_conv1_planes(request.param)
    yield request.param
    _lib.lib().drq_set_conv1_planes(prev)


@pytest.mark.parametrize("N,cin", [(3, 9), (40, 9), (5, 3), (2, 10)])
def test_conv1_bf16_fwd_and_wgrad(dev, N, cin, conv1_mode):
    """conv1 with fused integer-shift augmentation + normalisation, tensor-core path.  The kernel feeds
    the exact pixels (x - 128 in bf16) and bf16 weights, so the reference is fp64 math on exact inputs,
    bf16-rounded weights and (wgrad) the bf16-rounded gradient."""
    from drqv2_b200 import _lib
    from oracle import drq_oracle as O
    g = torch.Generator().manual_seed(N)
    obs = torch.randint(0, 256, (N, cin, 84, 84), dtype=torch.uint8, generator=g)
    shift = torch.randint(0, 9, (N, 2), dtype=torch.int32, generator=g)
    w = ((torch.rand(32, cin, 3, 3, generator=g) - 0.5) * 0.3).to(dev)
    b = ((torch.rand(32, generator=g) - 0.5) * 0.1).to(dev)
    wp = torch.zeros(_lib.lib().drq_conv1_w_packed_elems(), dtype=torch.bfloat16, device=dev)
    _lib.call("drq_pack_conv1_w_bf16", w.data_ptr(), b.data_ptr(), wp.data_ptr(), cin, _stream())
    out = torch.zeros(_lib.lib().drq_wb_elems(N), dtype=torch.bfloat16, device=dev)
    obs_d, shift_d = obs.to(dev), shift.to(dev)
    _lib.call("drq_conv1_fwd_bf16", obs_d.data_ptr(), shift_d.data_ptr(), wp.data_ptr(), out.data_ptr(), N, cin, 4, _stream())
    torch.cuda.synchronize()
    w16 = _bf(w).double()
    x = (O.random_shift_exact(obs.float(), shift).double() / 255.0 - 0.5).to(dev)
    want = torch.relu(torch.nn.functional.conv2d(x, w16, b.double(), stride=2))
    got = nchw_from_wb(out.view(4, -1, 8), N, 41, 41).double()
    assert (got - want).abs().max().item() <= 2 ** -8 * want.abs().max().item() + 1e-6
    # no augmentation (act path): shift = NULL == identity
    _lib.call("drq_conv1_fwd_bf16", obs_d.data_ptr(), None, wp.data_ptr(), out.data_ptr(), N, cin, 4, _stream())
    torch.cuda.synchronize()
    x0 = (obs.double() / 255.0 - 0.5).to(dev)
    want0 = torch.relu(torch.nn.functional.conv2d(x0, w16, b.double(), stride=2))
    assert (nchw_from_wb(out.view(4, -1, 8), N, 41, 41).double() - want0).abs().max().item() <= 2 ** -8 * want0.abs().max().item() + 1e-6
    # weight / bias gradient
    d = ((torch.rand(N, 32, 41, 41, generator=g) - 0.5) * 1e-2).to(dev)
    d_wb = wb_from_nchw(d)
    ws = torch.zeros(_lib.lib().drq_conv1_wgrad_bf16_ws_floats(), device=dev)
    dw, db = torch.zeros(32, cin, 3, 3, device=dev), torch.zeros(32, device=dev)
    _lib.call("drq_conv1_wgrad_bf16", obs_d.data_ptr(), shift_d.data_ptr(), d_wb.data_ptr(), ws.data_ptr(),
              dw.data_ptr(), db.data_ptr(), N, cin, 4, _stream())
    torch.cuda.synchronize()
    dr = _bf(d).double()
    want_w = torch.nn.grad.conv2d_weight(x, (32, cin, 3, 3), dr, stride=2)
    want_b = dr.sum(dim=(0, 2, 3))
    assert (dw.double() - want_w).abs().max().item() <= 1e-5 * want_w.abs().max().item() + 1e-7
    assert (db.double() - want_b).abs().max().item() <= 1e-5 * want_b.abs().max().item() + 1e-7


# ----------------------------------------------------------------------------- full update, bf16 mode
SCHED = "linear(1.0,0.1,100000)"


def _make_agent(A, Fd, H, lr, params, mode, use_graph=False):
    from drqv2_b200 import DrQV2Agent
    agent = DrQV2Agent((9, 84, 84), (A,), "cuda", lr, Fd, H, 0.01, 2000, 2, SCHED, 0.3, True,
                       use_cuda_graph=use_graph, seed=5, mode=mode)
    for net in ("encoder", "actor", "critic", "critic_target"):
        getattr(agent, net).load_state_dict(params[net])
    return agent


def _run(agent, b, step):
    agent.inject_draws(b["shift_obs"], b["shift_next"], b["eps_critic"], b["eps_actor"])
    return agent.update(iter([(b["obs"], b["action"], b["reward"], b["discount"], b["next_obs"])]), step)


# Stated bf16-mode tolerances (vs the fp64 oracle, one update from identical parameters and draws):
#   Q values / TD target / losses : 2e-2 relative (+1e-3 absolute)      (measured 3e-4 .. 8e-3)
#   gradients, per tensor, rel-L2  : twin-Q MLPs 0.15, critic trunk 0.2, encoder / actor 0.3 (B <= 16 here)
#                                    (measured 3e-3 .. 0.15 at B=16..192; cosine similarity >= 0.95 everywhere)
#   post-update parameters         : |delta| <= 2.5 lr per step (Adam's first steps are sign-like)
# Why gradients are this loose while every kernel is checked to 2^-8 above: an activation stored in
# bf16 carries ~0.3 % relative error, so ~0.3 % of the ReLU units (and a few min(Q1,Q2) selections)
# sit on the other side of zero than in fp64.  A flipped unit contributes its whole gradient as
# "error": rel-L2 ~ sqrt(flipped fraction) ~ 5 % per ReLU layer, accumulating down the backward
# chain (Q.4 0.3 % -> Q.0 3 % -> trunk 4 % -> conv4 6 % -> conv1 15 %, tools/bf16_error_scan*.py).
# It does not shrink with finer accumulation; it is the gradient of a slightly perturbed network.
BF16_METRIC_RTOL = 2e-2


def _bf16_grad_tol(net, name):
    if net == "critic" and name.startswith("Q"):
        return 0.15
    if net == "critic":
        return 0.2
    return 0.3


@pytest.mark.parametrize("case", [dict(B=16, A=6, F=50, H=256, lr=1e-4), dict(B=6, A=21, F=100, H=128, lr=8e-5)])
def test_update_bf16_matches_oracle(dev, case, capsys):
    """Stage-wise, as SURVEY §8c prescribes: with lr = 0 the optimiser steps are no-ops, so the
    actor stage sees the same critic as the oracle and every loss / Q / gradient is comparable at
    bf16 precision; with the real lr the post-update parameters are bounded by Adam's step size."""
    from oracle import drq_oracle as O
    from tests.helpers import rel_l2
    torch.set_num_threads(8)
    A, Fd, H, B, lr = case["A"], case["F"], case["H"], case["B"], case["lr"]
    params = O.synthetic_params(9, A, Fd, H, seed=4)
    b = O.synthetic_batch(B, A, seed=10)
    args = (b["obs"], b["action"], b["reward"], b["discount"], b["next_obs"], 0, b["shift_obs"],
            b["shift_next"], b["eps_critic"], b["eps_actor"])
    # --- lr = 0: all stages
    agent = _make_agent(A, Fd, H, 0.0, params, "bf16")
    o64 = O.OracleAgent(params, 0.0, 0.01, SCHED, 0.3, dtype=torch.float64)
    m, m64 = _run(agent, b, 0), o64.update(*args)
    report = {}
    for k in m64:
        report[k] = abs(m[k] - m64[k]) / (abs(m64[k]) + 1e-12)
        assert abs(m[k] - m64[k]) <= BF16_METRIC_RTOL * abs(m64[k]) + 1e-3, (k, m[k], m64[k])
    for net in ("encoder", "critic", "actor"):
        for name, p in getattr(agent, net).named_parameters():
            err = rel_l2(p.grad.cpu().numpy(), o64.grads[net][name].numpy())
            report[f"{net}.{name}"] = err
    with capsys.disabled():
        print("\nbf16 update errors vs fp64 oracle:", {k: float(f"{v:.2e}") for k, v in report.items()})
    for net in ("encoder", "critic", "actor"):
        for name, p in getattr(agent, net).named_parameters():
            assert report[f"{net}.{name}"] <= _bf16_grad_tol(net, name), (net, name, report[f"{net}.{name}"])
            g, w = p.grad.cpu().double().flatten(), o64.grads[net][name].flatten()
            cos = float(torch.dot(g, w) / (g.norm() * w.norm() + 1e-300))
            assert cos >= 0.95, (net, name, cos)
    # --- real lr: critic-side metrics and the parameter step bound
    agent = _make_agent(A, Fd, H, lr, params, "bf16")
    o64 = O.OracleAgent(params, lr, 0.01, SCHED, 0.3, dtype=torch.float64)
    m, m64 = _run(agent, b, 0), o64.update(*args)
    for k in ("batch_reward", "critic_target_q", "critic_q1", "critic_q2", "critic_loss"):
        assert abs(m[k] - m64[k]) <= BF16_METRIC_RTOL * abs(m64[k]) + 1e-3, (k, m[k], m64[k])
    for net in ("encoder", "critic", "actor", "critic_target"):
        for name, p in getattr(agent, net).named_parameters():
            assert (p.detach().cpu().double() - o64.p[net][name]).abs().max().item() <= 2.5 * lr, (net, name)


def test_bf16_graph_equals_eager_and_act(dev):
    from oracle import drq_oracle as O
    A, Fd, H, B = 6, 50, 128, 8
    params = O.synthetic_params(9, A, Fd, H, seed=6)
    eager = _make_agent(A, Fd, H, 1e-4, params, "bf16", use_graph=False)
    graph = _make_agent(A, Fd, H, 1e-4, params, "bf16", use_graph=True)
    for s in range(4):
        b = O.synthetic_batch(B, A, seed=50 + s)
        _run(eager, b, 2 * s)
        _run(graph, b, 2 * s)
    torch.cuda.synchronize()
    for net in ("encoder", "actor", "critic", "critic_target"):
        for (n1, p1), (n2, p2) in zip(getattr(eager, net).named_parameters(), getattr(graph, net).named_parameters()):
            assert torch.equal(p1, p2), (net, n1)
    o64 = O.OracleAgent(params, 1e-4, 0.01, SCHED, 0.3, dtype=torch.float64)
    fresh = _make_agent(A, Fd, H, 1e-4, params, "bf16", use_graph=True)
    obs = O.synthetic_batch(2, A, seed=5)["obs"]
    for i in range(2):
        a = fresh.act(obs[i].numpy(), 5000, True)
        want = o64.act(obs[i], 5000, True)[0].numpy()
        assert np.abs(a - want).max() < 2e-2


@pytest.mark.gpu
@pytest.mark.parametrize("A,Fd,H", [(6, 50, 1024), (1, 20, 256), (21, 50, 320)])
def test_fused_optimiser_step_equals_adam_then_pack(dev, A, Fd, H):
    """drq_adam_pack_step (Adam / soft update + bf16 operand refresh in one launch) against
    drq_adam_step / drq_adam_ema_step followed by drq_pack_multi: every fp32 arena and every packed
    operand bit for bit (torch.optim.Adam drqv2.py:148-150, utils.soft_update_params utils.py:42-45)."""
    import math
    from oracle import drq_oracle as O
    from drqv2_b200._lib import call
    F32 = 4
    params = O.synthetic_params(9, A, Fd, H, seed=3)
    agent = _make_agent(A, Fd, H, 1e-3, params, "bf16")
    st, a = agent._bf16, agent._arena
    s = torch.cuda.current_stream().cuda_stream
    g = torch.Generator(device="cuda").manual_seed(5)
    a.grads.copy_(torch.randn(a.total, device="cuda", generator=g) * 1e-2)
    a.exp_avg.copy_(torch.randn(a.total, device="cuda", generator=g) * 1e-2)
    a.exp_avg_sq.copy_(torch.rand(a.total, device="cuda", generator=g) * 1e-4)
    a.target.copy_(torch.randn(a.target.numel(), device="cuda", generator=g) * 0.1)
    live = torch.zeros(a.total, device="cuda")  # the arena's alignment padding carries no gradient / state
    for net in ("encoder", "critic", "actor"):
        for pname, prm in getattr(agent, net).named_parameters():
            live[a.offsets[net][pname]:a.offsets[net][pname] + prm.numel()] = 1
    for x in (a.grads, a.exp_avg, a.exp_avg_sq):
        x.mul_(live)
    c0, cn = a.seg["critic"][0], a.seg["critic"][2]
    a.target.mul_(live[c0:c0 + cn])
    t = 7
    for o in (0, 16, 24):                       # critic_opt / encoder_opt / actor_opt scalars of one slot (drqv2.SCAL_OFF)
        agent._scal_dev[o:o + 6] = torch.tensor([0.1, 0.999, 0.001, math.sqrt(1 - 0.999 ** t), 1e-8, -1e-3 / (1 - 0.9 ** t)])
    arenas = (a.params, a.exp_avg, a.exp_avg_sq, a.target)
    packed = [st.trunk.buf, st.q0.buf, st.q2.buf, st.p0.buf, st.p2.buf, st.p4.buf, st.conv1_w] + st.conv_wf + st.conv_wd
    snap = [x.clone() for x in arenas]

    def run(fused):
        for x, y in zip(arenas, snap):
            x.copy_(y)
        for x in packed:
            x.fill_(-7.0)                       # stale operands: every live element must be rewritten
        st.repack_all()
        if fused:
            st.step_critic_encoder()
            st.step_actor_target()
        else:
            off, n = a.seg["encoder"][0], a.seg["encoder"][2] + a.seg["critic"][2]
            call("drq_adam_step", a.params.data_ptr() + F32 * off, a.grads.data_ptr() + F32 * off,
                 a.exp_avg.data_ptr() + F32 * off, a.exp_avg_sq.data_ptr() + F32 * off, n, agent._scal_dev.data_ptr(), s)
            st.repack_critic_encoder()
            off, n = a.seg["actor"][0], a.seg["actor"][2]
            coff, cn = a.seg["critic"][0], a.seg["critic"][2]
            call("drq_adam_ema_step", a.params.data_ptr() + F32 * off, a.grads.data_ptr() + F32 * off,
                 a.exp_avg.data_ptr() + F32 * off, a.exp_avg_sq.data_ptr() + F32 * off, n, agent._scal_dev.data_ptr(),
                 a.params.data_ptr() + F32 * coff, a.target.data_ptr(), cn, 0.01, 0.99, s)
            st.repack_actor_target()
        torch.cuda.synchronize()
        return [x.clone() for x in arenas], [x.clone() for x in packed]

    agent.critic_target_tau = 0.01
    ref_a, ref_p = run(False)
    new_a, new_p = run(True)
    assert not torch.equal(ref_a[0], snap[0]) and not torch.equal(ref_a[3], snap[3])
    for name, x, y in zip(("params", "exp_avg", "exp_avg_sq", "target"), ref_a, new_a):
        assert torch.equal(x, y), (name, int((x != y).sum()))
    for i, (x, y) in enumerate(zip(ref_p, new_p)):
        assert torch.equal(x.view(torch.int16), y.view(torch.int16)), (i, int((x.view(torch.int16) != y.view(torch.int16)).sum()))


@pytest.mark.parametrize("M,A,H", [(512, 6, 1024), (70, 21, 256), (1, 12, 1024)])
def test_policy_head_matches_linear_tanh_sample(dev, M, A, H):
    """drq_policy_head_fwd_bf16: Linear(hidden, A) on bf16 operands (drqv2.py:81), tanh (:89) and the clipped
    TruncatedNormal samples of two row ranges (utils.py:117-126) in one launch, against the oracle's functions."""
    import ctypes as C
    from drqv2_b200 import _lib
    from drqv2_b200._bf16 import TB, PolicySample
    from oracle import drq_oracle as O
    g = torch.Generator().manual_seed(M + A)
    p2 = (torch.rand(M, H, generator=g) - 0.3).clamp_min(0).to(dev)
    w4 = ((torch.rand(A, H, generator=g) - 0.5) * 0.1).to(dev)
    b4 = ((torch.rand(A, generator=g) - 0.5) * 0.1).to(dev)
    p2b = TB(M, H, dev)
    p2b.load(p2)
    mu_pre = torch.zeros(M, A, device=dev)
    std = torch.tensor([0.4], device=dev)
    ticket = torch.zeros(1 + (M + 7) // 8, dtype=torch.int32, device=dev)       # block ticket + per-block log-prob sums
    n0 = M // 2
    rows = [(0, n0), (n0, M - n0)] if n0 else [(0, M)]
    eps = [torch.randn(r, A, generator=g).to(dev) for _, r in rows]
    act = [torch.zeros(r, A + 3, device=dev) for _, r in rows]          # written at a column offset (ld_a = A + 3)
    mu_out = torch.zeros(rows[-1][1], A, device=dev)
    metrics = torch.zeros(2, device=dev)
    xb = TB(rows[-1][1], 16 + A, dev)
    jobs = [PolicySample(r0, r, eps[i].data_ptr(), act[i].data_ptr() + 4 * 3, A + 3, None, None, None, 0, 0, 0) for i, (r0, r) in enumerate(rows)]
    jobs[-1] = PolicySample(rows[-1][0], rows[-1][1], eps[-1].data_ptr(), act[-1].data_ptr() + 4 * 3, A + 3, mu_out.data_ptr(),
                            metrics.data_ptr(), xb.ptr(), xb.units, 16, 0)
    arr = (PolicySample * len(jobs))(*jobs)
    for rep in range(2):                                 # twice: the block-ticket counter resets itself
        _lib.call("drq_policy_head_fwd_bf16", p2b.ptr(), p2b.units, w4.data_ptr(), b4.data_ptr(), mu_pre.data_ptr(), M, H, A,
                  arr, len(jobs), std.data_ptr(), 0.3, ticket.data_ptr(), _stream())
        torch.cuda.synchronize()
        assert int(ticket[0]) == 0
    want_pre = _bf(p2).double() @ _bf(w4).double().T + b4.double()
    assert (mu_pre.double() - want_pre).abs().max().item() <= 2e-5 * max(1.0, want_pre.abs().max().item())
    for i, (r0, r) in enumerate(rows):
        mu = torch.tanh(mu_pre[r0:r0 + r].cpu())
        want = O.truncated_normal_sample(mu, 0.4, eps[i].cpu(), 0.3)
        assert (act[i][:, 3:].cpu() - want).abs().max().item() <= 1e-6
    r0, r = rows[-1]
    mu = torch.tanh(mu_pre[r0:r0 + r].cpu())
    assert (mu_out.cpu() - mu).abs().max().item() <= 1e-6
    assert torch.equal(xb.dense()[:, 16:16 + A].cpu(), act[-1][:, 3:].cpu().to(torch.bfloat16).float())
    a = act[-1][:, 3:].cpu().double()
    lp = (-((a - mu.double()) ** 2) / (2 * 0.4 ** 2) - np.log(0.4) - np.log(np.sqrt(2 * np.pi))).sum(-1).mean().item()
    assert abs(float(metrics[0]) - lp) <= 1e-4 * abs(lp) + 1e-5
    assert abs(float(metrics[1]) - A * (0.5 + 0.5 * np.log(2 * np.pi) + np.log(0.4))) <= 1e-5
