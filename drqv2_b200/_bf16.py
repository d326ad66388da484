"""bf16 tensor-core mode of the DrQ-v2 update: the same stream of stages as the fp32 path in
drqv2.py (reference drqv2.py:177-262) with every dense contraction on tcgen05 kernels.

Master parameters, gradients and Adam state stay fp32 in the reference layouts; this module
owns the derived bf16 operand copies (re-packed after every optimiser step) and the bf16
activation buffers:
  * encoder activations / gradients: "WB" layout (include/drqv2_b200.h),
  * everything the heads touch (features, hidden activations, their gradients, Linear weights):
    the feature-blocked "FB" layout X_fb[f/8][row][8] - one buffer is both the K-major operand
    (contraction over features) and the MN-major operand (contraction over rows) of the GEMMs.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import (PLB, REPR_DIM, TEPI_F32, TEPI_MASK_BF16, TEPI_RELU_BF16, TEPI_TRUNK_DGRAD,
                   TEPI_TRUNK_WGRAD, WB_SLACK, call)

F32, BF = 4, 2


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ceil(n, m):
    return (n + m - 1) // m * m


def splitk_for(m_rows, K=REPR_DIM, target_blocks=148):
    mblocks = (m_rows + 127) // 128
    s0 = max(1, min(target_blocks // mblocks, K // 128))
    chunk = -(-K // s0)
    chunk = -(-chunk // 128) * 128
    return -(-K // chunk)


class FB:
    """`batch` bf16 matrices [rows][feats] in the feature-blocked layout (zero initialised)."""

    def __init__(self, rows, feats, dev, batch=1, rpad=None):
        self.rows, self.feats, self.batch = rows, feats, batch
        self.units = _ceil(feats, 16) // 8
        self.rpad = rpad if rpad is not None else _ceil(rows, 128)
        assert self.rpad >= rows
        self.stride = self.units * self.rpad * 8          # elements between batch entries
        self.buf = torch.zeros(batch * self.stride, dtype=torch.bfloat16, device=dev)

    def ptr(self, z=0, row=0, feat=0):
        assert feat % 8 == 0
        return self.buf.data_ptr() + BF * (z * self.stride + ((feat // 8) * self.rpad + row) * 8)

    def dense(self, z=0):
        """[rows][feats] fp32 view for tests."""
        x = self.buf[z * self.stride:(z + 1) * self.stride].view(self.units, self.rpad, 8)
        return x.permute(1, 0, 2).reshape(self.rpad, self.units * 8)[:self.rows, :self.feats].float()


def gemm(A, rpad_a, a_mn, B, rpad_b, b_mn, C, ldc, M, N, K, epi, bias=0, mask=0, rpad_mask=0, acc=0, batch=1,
         bs=(0, 0, 0, 0, 0), splitk=1, bn=None, n_store=0):
    if bn is None:
        bn = 32 if N <= 32 else 64
    call("drq_gemm_bf16", A, rpad_a, a_mn, B, rpad_b, b_mn, C, ldc, n_store, bias or None, mask or None, rpad_mask,
         M, N, K, epi, acc, batch, bs[0], bs[1], bs[2], bs[3], bs[4], splitk, bn, _stream())


class PackedNet:
    """FB bf16 copies of one network's Linear weights.  entries: name -> (fp32 offset in the
    net's parameter segment, rows, cols, heads, head stride in floats); 'trunk' is packed in
    the NHWC feature order of the bf16 encoder output."""

    def __init__(self, entries, dev):
        self.w, self.src = {}, entries
        for name, (_, rows, cols, heads, _) in entries.items():
            # rows padded for both roles: K-major B tiles (<= 128 rows per copy) and MN-major K chunks (64)
            self.w[name] = FB(rows, cols, dev, batch=heads, rpad=_ceil(rows, 128) if rows > 64 else 64)

    def ptr(self, name, z=0):
        return self.w[name].ptr(z)

    def repack(self, src_ptr):
        s = _stream()
        for name, (off, rows, cols, heads, hs) in self.src.items():
            fb = self.w[name]
            for z in range(heads):
                src = src_ptr + F32 * (off + z * hs)
                if name == "trunk":
                    call("drq_pack_trunk_fb", src, fb.ptr(z), rows, fb.rpad, s)
                else:
                    call("drq_pack_linear_fb", src, fb.ptr(z), rows, cols, fb.rpad, s)


class Bf16State:
    """Packed weights of an agent (independent of the batch size)."""

    def __init__(self, agent):
        dev = agent._dev
        a = agent._arena
        A, Fd, H = agent.action_dim, agent.feature_dim, agent.hidden_dim
        co = a.offsets["critic"]
        base_c = a.seg["critic"][0]
        qs = agent._q_strides()
        crit = {"trunk": (co["trunk.0.weight"] - base_c, Fd, REPR_DIM, 1, 0),
                "Q.0": (co["Q1.0.weight"] - base_c, H, Fd + A, 2, qs),
                "Q.2": (co["Q1.2.weight"] - base_c, H, H, 2, qs)}
        assert co["Q2.0.weight"] - co["Q1.0.weight"] == qs and co["Q2.2.weight"] - co["Q1.2.weight"] == qs
        self.critic = PackedNet(crit, dev)
        self.target = PackedNet(crit, dev)
        ao = a.offsets["actor"]
        base_a = a.seg["actor"][0]
        self.actor = PackedNet({
            "trunk": (ao["trunk.0.weight"] - base_a, Fd, REPR_DIM, 1, 0),
            "policy.0": (ao["policy.0.weight"] - base_a, H, Fd, 1, 0),
            "policy.2": (ao["policy.2.weight"] - base_a, H, H, 1, 0),
            "policy.4": (ao["policy.4.weight"] - base_a, A, H, 1, 0),
        }, dev)
        # encoder: conv1 packed + (fwd, dgrad) operands of conv2..4
        self.conv1_w = torch.zeros(12 * 32 * 8, dtype=torch.bfloat16, device=dev)
        self.conv_wf = [torch.zeros(36 * 32 * 8, dtype=torch.bfloat16, device=dev) for _ in range(3)]
        self.conv_wd = [torch.zeros(36 * 32 * 8, dtype=torch.bfloat16, device=dev) for _ in range(3)]
        self.agent = agent

    def repack_encoder(self):
        ag, s = self.agent, _stream()
        call("drq_pack_conv1_w_bf16", ag._p("encoder", "convnet.0.weight"), self.conv1_w.data_ptr(), ag.obs_shape[0], s)
        for i, k in enumerate((2, 4, 6)):
            call("drq_pack_conv_w_bf16", ag._p("encoder", f"convnet.{k}.weight"), self.conv_wf[i].data_ptr(),
                 self.conv_wd[i].data_ptr(), s)

    def repack_critic(self):
        self.critic.repack(self.agent._arena.ptr("params", "critic"))

    def repack_actor(self):
        self.actor.repack(self.agent._arena.ptr("params", "actor"))

    def repack_target(self):
        self.target.repack(self.agent._arena.target.data_ptr())

    def repack_all(self):
        self.repack_encoder()
        self.repack_critic()
        self.repack_actor()
        self.repack_target()


class Bf16Workspace:
    def __init__(self, B, A, Fd, H, st, dev):
        L = _lib.lib()
        zb = lambda *s: torch.zeros(*s, dtype=torch.bfloat16, device=dev)
        zf = lambda *s: torch.zeros(*s, device=dev)
        NB = 2 * B
        self.B = B
        self.acts = [zb(L.drq_wb_elems(NB)) for _ in range(3)]      # conv1..3 outputs, [obs | next]
        self.cs_act = NB * PLB + WB_SLACK
        self.feat = FB(NB, REPR_DIM, dev)                            # rows [obs | next], NHWC feature order
        self.dpre = [zb(L.drq_wb_elems(B)) for _ in range(4)]
        self.cs_d = B * PLB + WB_SLACK
        self.wg_ws = zf(max(L.drq_conv_wgrad_bf16_ws_floats(), L.drq_conv1_wgrad_bf16_ws_floats()))
        self.S = splitk_for(B)
        self.partial = zf(self.S * B * Fd)
        self.xT, self.xC, self.xA = FB(B, Fd + A, dev), FB(B, Fd + A, dev), FB(B, Fd + A, dev)
        self.hA = FB(B, Fd, dev)
        self.p1, self.p2 = FB(B, H, dev), FB(B, H, dev)
        self.c1, self.c2 = FB(B, H, dev, batch=2), FB(B, H, dev, batch=2)
        self.dc1, self.dc2 = FB(B, H, dev, batch=2), FB(B, H, dev, batch=2)
        self.dp1, self.dp2 = FB(B, H, dev), FB(B, H, dev)
        self.dz = FB(B, Fd, dev)
        self.dmu = FB(B, A, dev)
        self.dxf = zf(B, Fd + A)
        self.RP = self.p1.rpad


def encode(agent, ws, bw):
    """conv1 (u8 + aug + normalise fused) .. conv4 on tensor cores; features FB bf16 (NHWC order)."""
    st, B, s = agent._bf16, ws.B, _stream()
    be = lambda i: agent._p("encoder", f"convnet.{i}.bias")
    acts = [a.data_ptr() for a in bw.acts]
    call("drq_conv1_fwd_bf16", ws.obs.data_ptr(), ws.shift.data_ptr(), st.conv1_w.data_ptr(), be(0), acts[0],
         2 * B, agent.obs_shape[0], agent.aug.pad, s)
    call("drq_conv3x3_fwd_bf16", acts[0], st.conv_wf[0].data_ptr(), be(2), acts[1], 2 * B, 39, 0, 0, s)
    call("drq_conv3x3_fwd_bf16", acts[1], st.conv_wf[1].data_ptr(), be(4), acts[2], 2 * B, 37, 0, 0, s)
    call("drq_conv3x3_fwd_bf16", acts[2], st.conv_wf[2].data_ptr(), be(6), bw.feat.ptr(), 2 * B, 35, 2, bw.feat.rpad, s)


def trunk_fwd(agent, partial, S, feat, row0, B, pk, bias, gamma, beta, h_f32, ld_h, h_fb, xhat=0, rstd=0):
    """Linear(39200->F) as a split-K tensor-core GEMM + fused reduce/bias/LayerNorm/tanh."""
    Fd = agent.feature_dim
    wt = pk.w["trunk"]
    gemm(feat.ptr(row=row0), feat.rpad, 0, wt.ptr(), wt.rpad, 0, partial, Fd, B, Fd, REPR_DIM, TEPI_F32,
         splitk=S, bs=(0, 0, B * Fd, 0, 0), bn=64)
    call("drq_ln_tanh_fwd", partial, S, B * Fd, bias, gamma, beta, h_f32, ld_h, xhat or None, rstd or None,
         h_fb.ptr(), h_fb.rpad, B, Fd, 1e-5, _stream())


def actor_mlp_fwd(agent, hA, p1, p2, mu_pre, B):
    st = agent._bf16
    A, Fd, H = agent.action_dim, agent.feature_dim, agent.hidden_dim
    pa = lambda k: agent._p("actor", k)
    w0, w2, w4 = st.actor.w["policy.0"], st.actor.w["policy.2"], st.actor.w["policy.4"]
    gemm(hA.ptr(), hA.rpad, 0, w0.ptr(), w0.rpad, 0, p1.ptr(), p1.rpad, B, H, Fd, TEPI_RELU_BF16,
         bias=pa("policy.0.bias"))
    gemm(p1.ptr(), p1.rpad, 0, w2.ptr(), w2.rpad, 0, p2.ptr(), p2.rpad, B, H, H, TEPI_RELU_BF16,
         bias=pa("policy.2.bias"))
    gemm(p2.ptr(), p2.rpad, 0, w4.ptr(), w4.rpad, 0, mu_pre, A, B, A, H, TEPI_F32, bias=pa("policy.4.bias"), bn=32)


def twin_q_fwd(agent, bw, pk, x, pfn, q_out, B):
    """both Q heads per launch; pk = packed critic or target weights, pfn = fp32 param pointer fn."""
    A, Fd, H = agent.action_dim, agent.feature_dim, agent.hidden_dim
    qs_f = agent._q_strides()
    w0, w2 = pk.w["Q.0"], pk.w["Q.2"]
    c1, c2 = bw.c1, bw.c2
    gemm(x.ptr(), x.rpad, 0, w0.ptr(), w0.rpad, 0, c1.ptr(), c1.rpad, B, H, Fd + A, TEPI_RELU_BF16,
         bias=pfn("Q1.0.bias"), batch=2, bs=(0, w0.stride, c1.stride, qs_f, 0))
    gemm(c1.ptr(), c1.rpad, 0, w2.ptr(), w2.rpad, 0, c2.ptr(), c2.rpad, B, H, H, TEPI_RELU_BF16,
         bias=pfn("Q1.2.bias"), batch=2, bs=(c1.stride, w2.stride, c2.stride, qs_f, 0))
    call("drq_q_head_fwd_bf16", c2.ptr(), c2.rpad, c2.stride, pfn("Q1.4.weight"), pfn("Q1.4.bias"), q_out, B, H, 2,
         qs_f, _stream())


def critic_pass(agent, ws, bw):
    st, s = agent._bf16, _stream()
    B, A, Fd, H = ws.B, agent.action_dim, agent.feature_dim, agent.hidden_dim
    pc = lambda k: agent._p("critic", k)
    gc = lambda k: agent._g("critic", k)
    pa = lambda k: agent._p("actor", k)
    std_ptr = agent._scal_dev.data_ptr() + F32 * 8
    qs_f = agent._q_strides()
    feat = bw.feat
    part = bw.partial.data_ptr()
    # target: online actor on next features -> clipped sample
    trunk_fwd(agent, part, bw.S, feat, B, B, st.actor, pa("trunk.0.bias"), pa("trunk.1.weight"), pa("trunk.1.bias"),
              ws.hA.data_ptr(), Fd, bw.hA)
    actor_mlp_fwd(agent, bw.hA, bw.p1, bw.p2, ws.mu_pre.data_ptr(), B)
    call("drq_actor_sample", ws.mu_pre.data_ptr(), ws.eps_c.data_ptr(), std_ptr, float(agent.stddev_clip),
         ws.xT.data_ptr() + F32 * Fd, Fd + A, None, None, bw.xT.ptr(), bw.xT.rpad, Fd, B, A, s)
    trunk_fwd(agent, part, bw.S, feat, B, B, st.target, agent._t("trunk.0.bias"), agent._t("trunk.1.weight"),
              agent._t("trunk.1.bias"), ws.xT.data_ptr(), Fd + A, bw.xT)
    twin_q_fwd(agent, bw, st.target, bw.xT, agent._t, ws.tq.data_ptr(), B)
    # online critic
    trunk_fwd(agent, part, bw.S, feat, 0, B, st.critic, pc("trunk.0.bias"), pc("trunk.1.weight"), pc("trunk.1.bias"),
              ws.xC.data_ptr(), Fd + A, bw.xC, ws.xhatC.data_ptr(), ws.rstdC.data_ptr())
    call("drq_scatter_fb", ws.action.data_ptr(), A, bw.xC.ptr(), bw.xC.rpad, Fd, B, A, s)
    twin_q_fwd(agent, bw, st.critic, bw.xC, pc, ws.q.data_ptr(), B)
    q1, q2 = ws.q.data_ptr(), ws.q.data_ptr() + F32 * B
    call("drq_critic_loss", q1, q2, ws.tq.data_ptr(), ws.tq.data_ptr() + F32 * B, ws.reward.data_ptr(),
         ws.discount.data_ptr(), ws.dq.data_ptr(), ws.dq.data_ptr() + F32 * B, ws.target_q.data_ptr(),
         ws.metrics.data_ptr(), B, s)
    # backward through the Q heads
    c1, c2, dc1, dc2 = bw.c1, bw.c2, bw.dc1, bw.dc2
    w0, w2 = st.critic.w["Q.0"], st.critic.w["Q.2"]
    RP, HS = c1.rpad, c1.stride
    call("drq_q_head_bwd_bf16", ws.dq.data_ptr(), c2.ptr(), RP, HS, pc("Q1.4.weight"), dc2.ptr(),
         gc("Q1.4.weight"), gc("Q1.4.bias"), B, H, 2, qs_f, s)
    gemm(dc2.ptr(), RP, 1, c1.ptr(), RP, 1, gc("Q1.2.weight"), H, H, H, B, TEPI_F32, batch=2,
         bs=(HS, HS, qs_f, 0, 0), bn=128)
    call("drq_colsum_fb", dc2.ptr(), RP, gc("Q1.2.bias"), B, H, 2, HS, qs_f, s)
    gemm(dc2.ptr(), RP, 0, w2.ptr(), w2.rpad, 1, dc1.ptr(), RP, B, H, H, TEPI_MASK_BF16, mask=c1.ptr(), rpad_mask=RP,
         batch=2, bs=(HS, w2.stride, HS, 0, HS))
    gemm(dc1.ptr(), RP, 1, bw.xC.ptr(), bw.xC.rpad, 1, gc("Q1.0.weight"), Fd + A, H, Fd + A, B, TEPI_F32, batch=2,
         bs=(HS, 0, qs_f, 0, 0))
    call("drq_colsum_fb", dc1.ptr(), RP, gc("Q1.0.bias"), B, H, 2, HS, qs_f, s)
    gemm(dc1.ptr(0), RP, 0, w0.ptr(0), w0.rpad, 1, bw.dxf.data_ptr(), Fd + A, B, Fd, H, TEPI_F32)
    gemm(dc1.ptr(1), RP, 0, w0.ptr(1), w0.rpad, 1, bw.dxf.data_ptr(), Fd + A, B, Fd, H, TEPI_F32, acc=1)
    # trunk backward
    call("drq_ln_tanh_bwd", bw.dxf.data_ptr(), Fd + A, ws.xC.data_ptr(), Fd + A, ws.xhatC.data_ptr(),
         ws.rstdC.data_ptr(), pc("trunk.1.weight"), ws.dz.data_ptr(), gc("trunk.1.weight"), gc("trunk.1.bias"),
         bw.dz.ptr(), bw.dz.rpad, B, Fd, s)
    gemm(bw.dz.ptr(), bw.dz.rpad, 1, feat.ptr(), feat.rpad, 1, gc("trunk.0.weight"), REPR_DIM, Fd, REPR_DIM, B,
         TEPI_TRUNK_WGRAD, bn=128)
    call("drq_colsum_f32", ws.dz.data_ptr(), Fd, gc("trunk.0.bias"), B, Fd, 1, 0, 0, s)
    # encoder backward
    d = [t.data_ptr() for t in bw.dpre]
    acts = [t.data_ptr() for t in bw.acts]
    ge = lambda k: agent._g("encoder", k)
    wt = st.critic.w["trunk"]
    gemm(bw.dz.ptr(), bw.dz.rpad, 0, wt.ptr(), wt.rpad, 1, d[3], bw.cs_d, B, REPR_DIM, Fd, TEPI_TRUNK_DGRAD,
         mask=feat.ptr(), rpad_mask=feat.rpad, bn=128)
    wsp = bw.wg_ws.data_ptr()
    for layer, hout in ((3, 35), (2, 37), (1, 39)):
        k = 2 * layer
        call("drq_conv3x3_wgrad_bf16", acts[layer - 1], 2 * B, d[layer], wsp, ge(f"convnet.{k}.weight"),
             ge(f"convnet.{k}.bias"), B, hout, s)
        call("drq_conv3x3_dgrad_bf16", d[layer], st.conv_wd[layer - 1].data_ptr(), acts[layer - 1], 2 * B,
             d[layer - 1], B, hout, s)
    call("drq_conv1_wgrad_bf16", ws.obs.data_ptr(), ws.shift.data_ptr(), d[0], wsp, ge("convnet.0.weight"),
         ge("convnet.0.bias"), B, agent.obs_shape[0], agent.aug.pad, s)
    # critic_opt.step(); encoder_opt.step(); refresh their bf16 operand copies
    a = agent._arena
    off, n = a.seg["encoder"][0], a.seg["encoder"][2] + a.seg["critic"][2]
    call("drq_adam_step", a.params.data_ptr() + F32 * off, a.grads.data_ptr() + F32 * off,
         a.exp_avg.data_ptr() + F32 * off, a.exp_avg_sq.data_ptr() + F32 * off, n, agent._scal_dev.data_ptr(), s)
    st.repack_critic()
    st.repack_encoder()


def actor_pass(agent, ws, bw):
    st, s = agent._bf16, _stream()
    B, A, Fd, H = ws.B, agent.action_dim, agent.feature_dim, agent.hidden_dim
    pc = lambda k: agent._p("critic", k)
    pa = lambda k: agent._p("actor", k)
    ga = lambda k: agent._g("actor", k)
    std_ptr = agent._scal_dev.data_ptr() + F32 * 8
    qs_f = agent._q_strides()
    feat = bw.feat
    part = bw.partial.data_ptr()
    trunk_fwd(agent, part, bw.S, feat, 0, B, st.actor, pa("trunk.0.bias"), pa("trunk.1.weight"), pa("trunk.1.bias"),
              ws.hA.data_ptr(), Fd, bw.hA, ws.xhatA.data_ptr(), ws.rstdA.data_ptr())
    actor_mlp_fwd(agent, bw.hA, bw.p1, bw.p2, ws.mu_pre.data_ptr(), B)
    call("drq_actor_sample", ws.mu_pre.data_ptr(), ws.eps_a.data_ptr(), std_ptr, float(agent.stddev_clip),
         ws.xA.data_ptr() + F32 * Fd, Fd + A, ws.mu.data_ptr(), ws.metrics.data_ptr() + F32 * 6,
         bw.xA.ptr(), bw.xA.rpad, Fd, B, A, s)
    trunk_fwd(agent, part, bw.S, feat, 0, B, st.critic, pc("trunk.0.bias"), pc("trunk.1.weight"), pc("trunk.1.bias"),
              ws.xA.data_ptr(), Fd + A, bw.xA)
    twin_q_fwd(agent, bw, st.critic, bw.xA, pc, ws.q.data_ptr(), B)
    call("drq_actor_loss", ws.q.data_ptr(), ws.q.data_ptr() + F32 * B, ws.dq.data_ptr(), ws.dq.data_ptr() + F32 * B,
         ws.metrics.data_ptr() + F32 * 5, B, s)
    c1, c2, dc1, dc2 = bw.c1, bw.c2, bw.dc1, bw.dc2
    w0, w2 = st.critic.w["Q.0"], st.critic.w["Q.2"]
    RP, HS = c1.rpad, c1.stride
    call("drq_q_head_bwd_bf16", ws.dq.data_ptr(), c2.ptr(), RP, HS, pc("Q1.4.weight"), dc2.ptr(), None, None,
         B, H, 2, qs_f, s)
    gemm(dc2.ptr(), RP, 0, w2.ptr(), w2.rpad, 1, dc1.ptr(), RP, B, H, H, TEPI_MASK_BF16, mask=c1.ptr(), rpad_mask=RP,
         batch=2, bs=(HS, w2.stride, HS, 0, HS))
    gemm(dc1.ptr(0), RP, 0, w0.ptr(0), w0.rpad, 1, bw.dxf.data_ptr(), Fd + A, B, Fd + A, H, TEPI_F32)
    gemm(dc1.ptr(1), RP, 0, w0.ptr(1), w0.rpad, 1, bw.dxf.data_ptr(), Fd + A, B, Fd + A, H, TEPI_F32, acc=1)
    call("drq_actor_sample_bwd", bw.dxf.data_ptr() + F32 * Fd, Fd + A, ws.mu.data_ptr(), ws.dmu_pre.data_ptr(),
         bw.dmu.ptr(), bw.dmu.rpad, B, A, s)
    # actor MLP backward
    dmu, hA, p1, p2, dp1, dp2 = bw.dmu, bw.hA, bw.p1, bw.p2, bw.dp1, bw.dp2
    a0, a2, a4 = st.actor.w["policy.0"], st.actor.w["policy.2"], st.actor.w["policy.4"]
    gemm(dmu.ptr(), RP, 1, p2.ptr(), RP, 1, ga("policy.4.weight"), H, A, H, B, TEPI_F32, bn=128)
    call("drq_colsum_f32", ws.dmu_pre.data_ptr(), A, ga("policy.4.bias"), B, A, 1, 0, 0, s)
    gemm(dmu.ptr(), RP, 0, a4.ptr(), a4.rpad, 1, dp2.ptr(), RP, B, H, A, TEPI_MASK_BF16, mask=p2.ptr(), rpad_mask=RP)
    gemm(dp2.ptr(), RP, 1, p1.ptr(), RP, 1, ga("policy.2.weight"), H, H, H, B, TEPI_F32, bn=128)
    call("drq_colsum_fb", dp2.ptr(), RP, ga("policy.2.bias"), B, H, 1, 0, 0, s)
    gemm(dp2.ptr(), RP, 0, a2.ptr(), a2.rpad, 1, dp1.ptr(), RP, B, H, H, TEPI_MASK_BF16, mask=p1.ptr(), rpad_mask=RP)
    gemm(dp1.ptr(), RP, 1, hA.ptr(), RP, 1, ga("policy.0.weight"), Fd, H, Fd, B, TEPI_F32)
    call("drq_colsum_fb", dp1.ptr(), RP, ga("policy.0.bias"), B, H, 1, 0, 0, s)
    gemm(dp1.ptr(), RP, 0, a0.ptr(), a0.rpad, 1, ws.dhA.data_ptr(), Fd, B, Fd, H, TEPI_F32)
    call("drq_ln_tanh_bwd", ws.dhA.data_ptr(), Fd, ws.hA.data_ptr(), Fd, ws.xhatA.data_ptr(), ws.rstdA.data_ptr(),
         pa("trunk.1.weight"), ws.dz.data_ptr(), ga("trunk.1.weight"), ga("trunk.1.bias"), bw.dz.ptr(), bw.dz.rpad,
         B, Fd, s)
    gemm(bw.dz.ptr(), bw.dz.rpad, 1, feat.ptr(), feat.rpad, 1, ga("trunk.0.weight"), REPR_DIM, Fd, REPR_DIM, B,
         TEPI_TRUNK_WGRAD, bn=128)
    call("drq_colsum_f32", ws.dz.data_ptr(), Fd, ga("trunk.0.bias"), B, Fd, 1, 0, 0, s)
    a = agent._arena
    off, n = a.seg["actor"][0], a.seg["actor"][2]
    coff, cn = a.seg["critic"][0], a.seg["critic"][2]
    tau = float(agent.critic_target_tau)
    call("drq_adam_ema_step", a.params.data_ptr() + F32 * off, a.grads.data_ptr() + F32 * off,
         a.exp_avg.data_ptr() + F32 * off, a.exp_avg_sq.data_ptr() + F32 * off, n, agent._scal_dev.data_ptr(),
         a.params.data_ptr() + F32 * coff, a.target.data_ptr(), cn, tau, float(1 - tau), s)
    st.repack_actor()
    st.repack_target()


def act_workspace(agent, n, dev):
    A, Fd, H = agent.action_dim, agent.feature_dim, agent.hidden_dim
    nel = _lib.lib().drq_wb_elems(n)
    zb = lambda *s: torch.zeros(*s, dtype=torch.bfloat16, device=dev)
    S = splitk_for(n)
    return dict(acts_b=[zb(nel) for _ in range(3)], feat_b=FB(n, REPR_DIM, dev), S_b=S,
                partial_b=torch.zeros(S * n * Fd, device=dev), h_b=FB(n, Fd, dev), p1_b=FB(n, H, dev),
                p2_b=FB(n, H, dev))


def act_body(agent, w, n, sample):
    """encoder + actor for act() at batch n on the tensor-core path."""
    st, s = agent._bf16, _stream()
    A, Fd = agent.action_dim, agent.feature_dim
    if sample:
        call("drq_rng_normal_f32", agent._seed, agent._counter.data_ptr(), w["eps"].data_ptr(), n * A, s)
        call("drq_counter_advance", agent._counter.data_ptr(), s)
    be = lambda i: agent._p("encoder", f"convnet.{i}.bias")
    acts = [a.data_ptr() for a in w["acts_b"]]
    feat = w["feat_b"]
    call("drq_conv1_fwd_bf16", w["obs"].data_ptr(), None, st.conv1_w.data_ptr(), be(0), acts[0], n,
         agent.obs_shape[0], agent.aug.pad, s)
    call("drq_conv3x3_fwd_bf16", acts[0], st.conv_wf[0].data_ptr(), be(2), acts[1], n, 39, 0, 0, s)
    call("drq_conv3x3_fwd_bf16", acts[1], st.conv_wf[1].data_ptr(), be(4), acts[2], n, 37, 0, 0, s)
    call("drq_conv3x3_fwd_bf16", acts[2], st.conv_wf[2].data_ptr(), be(6), feat.ptr(), n, 35, 2, feat.rpad, s)
    pa = lambda k: agent._p("actor", k)
    trunk_fwd(agent, w["partial_b"].data_ptr(), w["S_b"], feat, 0, n, st.actor, pa("trunk.0.bias"),
              pa("trunk.1.weight"), pa("trunk.1.bias"), w["h"].data_ptr(), Fd, w["h_b"])
    actor_mlp_fwd(agent, w["h_b"], w["p1_b"], w["p2_b"], w["mu_pre"].data_ptr(), n)
    call("drq_actor_sample", w["mu_pre"].data_ptr(), w["eps"].data_ptr() if sample else None,
         agent._scal_dev.data_ptr() + F32 * 8, 0.0, w["out"].data_ptr(), A, None, None, None, 0, 0, n, A, s)
