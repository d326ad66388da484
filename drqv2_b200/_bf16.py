"""bf16 tensor-core mode of the DrQ-v2 update: the same stream of stages as the fp32 path in
drqv2.py (reference drqv2.py:177-262) with every dense contraction on tcgen05 kernels.

Master parameters, gradients and Adam state stay fp32 in the reference layouts; this module
owns the derived bf16 operand copies (re-packed after every optimiser step) and the bf16
activation buffers:
  * encoder activations / gradients: "WB" layout (include/drqv2_b200.h),
  * features: NHWC-compact [N][35*35][32] bf16 (the trunk weight copy is permuted to match),
  * head activations: row-major bf16 with row strides padded to 8 elements (zero padded).
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import (PLB, REPR_DIM, TEPI_F32, TEPI_MASK_BF16, TEPI_RELU_BF16, TEPI_TRUNK_DGRAD,
                   TEPI_TRUNK_WGRAD, WB_SLACK, call)

F32, BF = 4, 2


def _stream():
    return torch.cuda.current_stream().cuda_stream


def pad8(n):
    return (n + 7) // 8 * 8


def splitk_for(m_rows, K=REPR_DIM, target_blocks=148):
    mblocks = (m_rows + 127) // 128
    s0 = max(1, min(target_blocks // mblocks, K // 128))
    chunk = -(-K // s0)
    chunk = -(-chunk // 128) * 128
    return -(-K // chunk)


def gemm(A, lda, a_mn, B, ldb, b_mn, C, ldc, M, N, K, epi, bias=0, mask=0, ldmask=0, acc=0, batch=1,
         bs=(0, 0, 0, 0, 0), splitk=1, bn=None):
    if bn is None:
        bn = 32 if N <= 32 else 64
    call("drq_gemm_bf16", A, lda, a_mn, B, ldb, b_mn, C, ldc, bias or None, mask or None, ldmask, M, N, K, epi,
         acc, batch, bs[0], bs[1], bs[2], bs[3], bs[4], splitk, bn, _stream())


class PackedNet:
    """bf16 copies of one network's Linear weights + the device table that re-packs them."""

    def __init__(self, entries, dev):
        # entries: name -> (src_off_floats, rows, cols, ld, nhwc)
        self.off, rows_tbl, total, self.trunk = {}, [], 0, None
        for name, (src, rows, cols, ld, nhwc) in entries.items():
            self.off[name] = total
            if nhwc:
                self.trunk = (src, total, rows)        # packed by the transposing kernel
            else:
                rows_tbl.append([src, total, rows, cols, ld, nhwc])
            total += pad8(rows) * ld + 64          # row padding: MN-major reads of 8-row units stay in bounds
        self.buf = torch.zeros(total + 64, dtype=torch.bfloat16, device=dev)
        self.table = torch.tensor(rows_tbl, dtype=torch.int64, device=dev)
        self.n = len(rows_tbl)

    def ptr(self, name):
        return self.buf.data_ptr() + BF * self.off[name]

    def repack(self, src_ptr):
        call("drq_pack_table_bf16", src_ptr, self.buf.data_ptr(), self.table.data_ptr(), self.n, _stream())
        if self.trunk is not None:
            src, dst, rows = self.trunk
            call("drq_pack_trunk_bf16", src_ptr + F32 * src, self.buf.data_ptr() + BF * dst, rows, _stream())


class Bf16State:
    """Packed weights of an agent (independent of the batch size)."""

    def __init__(self, agent):
        dev = agent._dev
        a = agent._arena
        A, Fd, H = agent.action_dim, agent.feature_dim, agent.hidden_dim
        self.ldF, self.ldX, self.ldA = pad8(Fd), pad8(Fd + A), pad8(A)
        co = a.offsets["critic"]
        base_c = a.seg["critic"][0]
        crit = {"trunk": (co["trunk.0.weight"] - base_c, Fd, REPR_DIM, REPR_DIM, 1)}
        for q in ("Q1", "Q2"):
            crit[f"{q}.0"] = (co[f"{q}.0.weight"] - base_c, H, Fd + A, self.ldX, 0)
            crit[f"{q}.2"] = (co[f"{q}.2.weight"] - base_c, H, H, H, 0)
        self.critic = PackedNet(crit, dev)
        self.target = PackedNet(crit, dev)
        ao = a.offsets["actor"]
        base_a = a.seg["actor"][0]
        self.actor = PackedNet({
            "trunk": (ao["trunk.0.weight"] - base_a, Fd, REPR_DIM, REPR_DIM, 1),
            "policy.0": (ao["policy.0.weight"] - base_a, H, Fd, self.ldF, 0),
            "policy.2": (ao["policy.2.weight"] - base_a, H, H, H, 0),
            "policy.4": (ao["policy.4.weight"] - base_a, A, H, H, 0),
        }, dev)
        self.q_stride_b = self.critic.off["Q2.0"] - self.critic.off["Q1.0"]     # bf16 elements between heads
        assert self.critic.off["Q2.2"] - self.critic.off["Q1.2"] == self.q_stride_b
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
        self.feat = zb(NB, REPR_DIM)                                  # NHWC-compact
        self.dpre = [zb(L.drq_wb_elems(B)) for _ in range(4)]
        self.cs_d = B * PLB + WB_SLACK
        self.wg_ws = zf(max(L.drq_conv_wgrad_bf16_ws_floats(), L.drq_conv1_wgrad_bf16_ws_floats()))
        self.S = splitk_for(B)
        self.partial = zf(self.S * B * Fd)
        ldF, ldX, ldA = st.ldF, st.ldX, st.ldA
        self.xT, self.xC, self.xA = zb(B, ldX), zb(B, ldX), zb(B, ldX)
        self.hA = zb(B, ldF)
        self.p1, self.p2 = zb(B, H), zb(B, H)
        self.c1, self.c2 = zb(2, B, H), zb(2, B, H)
        self.dc1, self.dc2 = zb(2, B, H), zb(2, B, H)
        self.dp1, self.dp2 = zb(B, H), zb(B, H)
        self.dz = zb(B, ldF)
        self.dmu = zb(B, ldA)
        self.dxf = zf(B, Fd + A)


def encode(agent, ws, bw):
    """conv1 (u8 + aug + normalise fused) .. conv4 on tensor cores; features NHWC bf16."""
    st, B, s = agent._bf16, ws.B, _stream()
    be = lambda i: agent._p("encoder", f"convnet.{i}.bias")
    acts = [a.data_ptr() for a in bw.acts]
    call("drq_conv1_fwd_bf16", ws.obs.data_ptr(), ws.shift.data_ptr(), st.conv1_w.data_ptr(), be(0), acts[0],
         2 * B, agent.obs_shape[0], agent.aug.pad, s)
    call("drq_conv3x3_fwd_bf16", acts[0], st.conv_wf[0].data_ptr(), be(2), acts[1], 2 * B, 39, 0, s)
    call("drq_conv3x3_fwd_bf16", acts[1], st.conv_wf[1].data_ptr(), be(4), acts[2], 2 * B, 37, 0, s)
    call("drq_conv3x3_fwd_bf16", acts[2], st.conv_wf[2].data_ptr(), be(6), bw.feat.data_ptr(), 2 * B, 35, 1, s)


def trunk_fwd(agent, bw, feat_ptr, B, pk, bias, gamma, beta, h_f32, ld_h, h_bf, ld_hb, xhat=0, rstd=0):
    Fd = agent.feature_dim
    gemm(feat_ptr, REPR_DIM, 0, pk.ptr("trunk"), REPR_DIM, 0, bw.partial.data_ptr(), Fd, B, Fd, REPR_DIM, TEPI_F32,
         splitk=bw.S, bs=(0, 0, B * Fd, 0, 0), bn=64)
    call("drq_ln_tanh_fwd", bw.partial.data_ptr(), bw.S, B * Fd, bias, gamma, beta, h_f32, ld_h, xhat or None,
         rstd or None, h_bf, ld_hb, B, Fd, 1e-5, _stream())


def actor_mlp_fwd(agent, bw, B):
    st = agent._bf16
    A, Fd, H = agent.action_dim, agent.feature_dim, agent.hidden_dim
    pa = lambda k: agent._p("actor", k)
    ws = agent.workspace(B)
    gemm(bw.hA.data_ptr(), st.ldF, 0, st.actor.ptr("policy.0"), st.ldF, 0, bw.p1.data_ptr(), H, B, H, Fd,
         TEPI_RELU_BF16, bias=pa("policy.0.bias"))
    gemm(bw.p1.data_ptr(), H, 0, st.actor.ptr("policy.2"), H, 0, bw.p2.data_ptr(), H, B, H, H, TEPI_RELU_BF16,
         bias=pa("policy.2.bias"))
    gemm(bw.p2.data_ptr(), H, 0, st.actor.ptr("policy.4"), H, 0, ws.mu_pre.data_ptr(), A, B, A, H, TEPI_F32,
         bias=pa("policy.4.bias"), bn=32)


def twin_q_fwd(agent, bw, pk, x_ptr, pfn, q_out, B):
    """both Q heads per launch; pk = packed critic or target weights, pfn = fp32 param pointer fn."""
    st = agent._bf16
    A, Fd, H = agent.action_dim, agent.feature_dim, agent.hidden_dim
    qs_f, qs_b = agent._q_strides(), st.q_stride_b
    BH = B * H
    gemm(x_ptr, st.ldX, 0, pk.ptr("Q1.0"), st.ldX, 0, bw.c1.data_ptr(), H, B, H, Fd + A, TEPI_RELU_BF16,
         bias=pfn("Q1.0.bias"), batch=2, bs=(0, qs_b, BH, qs_f, 0))
    gemm(bw.c1.data_ptr(), H, 0, pk.ptr("Q1.2"), H, 0, bw.c2.data_ptr(), H, B, H, H, TEPI_RELU_BF16,
         bias=pfn("Q1.2.bias"), batch=2, bs=(BH, qs_b, BH, qs_f, 0))
    call("drq_q_head_fwd_bf16", bw.c2.data_ptr(), pfn("Q1.4.weight"), pfn("Q1.4.bias"), q_out, B, H, 2, qs_f,
         _stream())


def critic_pass(agent, ws, bw):
    st, s = agent._bf16, _stream()
    B, A, Fd, H = ws.B, agent.action_dim, agent.feature_dim, agent.hidden_dim
    pc = lambda k: agent._p("critic", k)
    gc = lambda k: agent._g("critic", k)
    pa = lambda k: agent._p("actor", k)
    std_ptr = agent._scal_dev.data_ptr() + F32 * 8
    qs_f, qs_b = agent._q_strides(), st.q_stride_b
    BH = B * H
    feat_o, feat_n = bw.feat.data_ptr(), bw.feat.data_ptr() + BF * B * REPR_DIM
    ldX, ldF = st.ldX, st.ldF
    # target: online actor on next features -> clipped sample
    trunk_fwd(agent, bw, feat_n, B, st.actor, pa("trunk.0.bias"), pa("trunk.1.weight"), pa("trunk.1.bias"),
              ws.hA.data_ptr(), Fd, bw.hA.data_ptr(), ldF)
    actor_mlp_fwd(agent, bw, B)
    call("drq_actor_sample", ws.mu_pre.data_ptr(), ws.eps_c.data_ptr(), std_ptr, float(agent.stddev_clip),
         ws.xT.data_ptr() + F32 * Fd, Fd + A, None, None, bw.xT.data_ptr() + BF * Fd, ldX, B, A, s)
    trunk_fwd(agent, bw, feat_n, B, st.target, agent._t("trunk.0.bias"), agent._t("trunk.1.weight"),
              agent._t("trunk.1.bias"), ws.xT.data_ptr(), Fd + A, bw.xT.data_ptr(), ldX)
    twin_q_fwd(agent, bw, st.target, bw.xT.data_ptr(), agent._t, ws.tq.data_ptr(), B)
    # online critic
    trunk_fwd(agent, bw, feat_o, B, st.critic, pc("trunk.0.bias"), pc("trunk.1.weight"), pc("trunk.1.bias"),
              ws.xC.data_ptr(), Fd + A, bw.xC.data_ptr(), ldX, ws.xhatC.data_ptr(), ws.rstdC.data_ptr())
    call("drq_copy2d_f32_bf16", ws.action.data_ptr(), A, bw.xC.data_ptr() + BF * Fd, ldX, B, A, s)
    twin_q_fwd(agent, bw, st.critic, bw.xC.data_ptr(), pc, ws.q.data_ptr(), B)
    q1, q2 = ws.q.data_ptr(), ws.q.data_ptr() + F32 * B
    call("drq_critic_loss", q1, q2, ws.tq.data_ptr(), ws.tq.data_ptr() + F32 * B, ws.reward.data_ptr(),
         ws.discount.data_ptr(), ws.dq.data_ptr(), ws.dq.data_ptr() + F32 * B, ws.target_q.data_ptr(),
         ws.metrics.data_ptr(), B, s)
    # backward through the Q heads
    c1, c2, dc1, dc2 = (t.data_ptr() for t in (bw.c1, bw.c2, bw.dc1, bw.dc2))
    call("drq_q_head_bwd_bf16", ws.dq.data_ptr(), c2, pc("Q1.4.weight"), dc2, gc("Q1.4.weight"), gc("Q1.4.bias"),
         B, H, 2, qs_f, s)
    gemm(dc2, H, 1, c1, H, 1, gc("Q1.2.weight"), H, H, H, B, TEPI_F32, batch=2, bs=(BH, BH, qs_f, 0, 0), bn=128)
    call("drq_colsum_bf16", dc2, H, gc("Q1.2.bias"), B, H, 2, BH, qs_f, s)
    gemm(dc2, H, 0, st.critic.ptr("Q1.2"), H, 1, dc1, H, B, H, H, TEPI_MASK_BF16, mask=c1, ldmask=H, batch=2,
         bs=(BH, qs_b, BH, 0, BH))
    gemm(dc1, H, 1, bw.xC.data_ptr(), ldX, 1, gc("Q1.0.weight"), Fd + A, H, Fd + A, B, TEPI_F32, batch=2,
         bs=(BH, 0, qs_f, 0, 0))
    call("drq_colsum_bf16", dc1, H, gc("Q1.0.bias"), B, H, 2, BH, qs_f, s)
    gemm(dc1, H, 0, st.critic.ptr("Q1.0"), ldX, 1, bw.dxf.data_ptr(), Fd + A, B, Fd, H, TEPI_F32)
    gemm(dc1 + BF * BH, H, 0, st.critic.ptr("Q2.0"), ldX, 1, bw.dxf.data_ptr(), Fd + A, B, Fd, H, TEPI_F32, acc=1)
    # trunk backward
    call("drq_ln_tanh_bwd", bw.dxf.data_ptr(), Fd + A, ws.xC.data_ptr(), Fd + A, ws.xhatC.data_ptr(),
         ws.rstdC.data_ptr(), pc("trunk.1.weight"), ws.dz.data_ptr(), gc("trunk.1.weight"), gc("trunk.1.bias"),
         bw.dz.data_ptr(), ldF, B, Fd, s)
    gemm(bw.dz.data_ptr(), ldF, 1, feat_o, REPR_DIM, 1, gc("trunk.0.weight"), REPR_DIM, Fd, REPR_DIM, B,
         TEPI_TRUNK_WGRAD, bn=128)
    call("drq_colsum_f32", ws.dz.data_ptr(), Fd, gc("trunk.0.bias"), B, Fd, 1, 0, 0, s)
    # encoder backward
    d = [t.data_ptr() for t in bw.dpre]
    acts = [t.data_ptr() for t in bw.acts]
    ge = lambda k: agent._g("encoder", k)
    gemm(bw.dz.data_ptr(), ldF, 0, st.critic.ptr("trunk"), REPR_DIM, 1, d[3], bw.cs_d, B, REPR_DIM, Fd,
         TEPI_TRUNK_DGRAD, mask=feat_o, ldmask=REPR_DIM, bn=128)
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
    qs_f, qs_b = agent._q_strides(), st.q_stride_b
    BH = B * H
    feat_o = bw.feat.data_ptr()
    ldX, ldF, ldA = st.ldX, st.ldF, st.ldA
    trunk_fwd(agent, bw, feat_o, B, st.actor, pa("trunk.0.bias"), pa("trunk.1.weight"), pa("trunk.1.bias"),
              ws.hA.data_ptr(), Fd, bw.hA.data_ptr(), ldF, ws.xhatA.data_ptr(), ws.rstdA.data_ptr())
    actor_mlp_fwd(agent, bw, B)
    call("drq_actor_sample", ws.mu_pre.data_ptr(), ws.eps_a.data_ptr(), std_ptr, float(agent.stddev_clip),
         ws.xA.data_ptr() + F32 * Fd, Fd + A, ws.mu.data_ptr(), ws.metrics.data_ptr() + F32 * 6,
         bw.xA.data_ptr() + BF * Fd, ldX, B, A, s)
    trunk_fwd(agent, bw, feat_o, B, st.critic, pc("trunk.0.bias"), pc("trunk.1.weight"), pc("trunk.1.bias"),
              ws.xA.data_ptr(), Fd + A, bw.xA.data_ptr(), ldX)
    twin_q_fwd(agent, bw, st.critic, bw.xA.data_ptr(), pc, ws.q.data_ptr(), B)
    call("drq_actor_loss", ws.q.data_ptr(), ws.q.data_ptr() + F32 * B, ws.dq.data_ptr(), ws.dq.data_ptr() + F32 * B,
         ws.metrics.data_ptr() + F32 * 5, B, s)
    c1, c2, dc1, dc2 = (t.data_ptr() for t in (bw.c1, bw.c2, bw.dc1, bw.dc2))
    call("drq_q_head_bwd_bf16", ws.dq.data_ptr(), c2, pc("Q1.4.weight"), dc2, None, None, B, H, 2, qs_f, s)
    gemm(dc2, H, 0, st.critic.ptr("Q1.2"), H, 1, dc1, H, B, H, H, TEPI_MASK_BF16, mask=c1, ldmask=H, batch=2,
         bs=(BH, qs_b, BH, 0, BH))
    gemm(dc1, H, 0, st.critic.ptr("Q1.0"), ldX, 1, bw.dxf.data_ptr(), Fd + A, B, Fd + A, H, TEPI_F32)
    gemm(dc1 + BF * BH, H, 0, st.critic.ptr("Q2.0"), ldX, 1, bw.dxf.data_ptr(), Fd + A, B, Fd + A, H, TEPI_F32, acc=1)
    call("drq_actor_sample_bwd", bw.dxf.data_ptr() + F32 * Fd, Fd + A, ws.mu.data_ptr(), ws.dmu_pre.data_ptr(),
         bw.dmu.data_ptr(), ldA, B, A, s)
    # actor MLP backward
    dmu, p1, p2, dp1, dp2 = (t.data_ptr() for t in (bw.dmu, bw.p1, bw.p2, bw.dp1, bw.dp2))
    gemm(dmu, ldA, 1, p2, H, 1, ga("policy.4.weight"), H, A, H, B, TEPI_F32, bn=128)
    call("drq_colsum_f32", ws.dmu_pre.data_ptr(), A, ga("policy.4.bias"), B, A, 1, 0, 0, s)
    gemm(dmu, ldA, 0, st.actor.ptr("policy.4"), H, 1, dp2, H, B, H, A, TEPI_MASK_BF16, mask=p2, ldmask=H)
    gemm(dp2, H, 1, p1, H, 1, ga("policy.2.weight"), H, H, H, B, TEPI_F32, bn=128)
    call("drq_colsum_bf16", dp2, H, ga("policy.2.bias"), B, H, 1, 0, 0, s)
    gemm(dp2, H, 0, st.actor.ptr("policy.2"), H, 1, dp1, H, B, H, H, TEPI_MASK_BF16, mask=p1, ldmask=H)
    gemm(dp1, H, 1, bw.hA.data_ptr(), ldF, 1, ga("policy.0.weight"), Fd, H, Fd, B, TEPI_F32)
    call("drq_colsum_bf16", dp1, H, ga("policy.0.bias"), B, H, 1, 0, 0, s)
    gemm(dp1, H, 0, st.actor.ptr("policy.0"), ldF, 1, ws.dhA.data_ptr(), Fd, B, Fd, H, TEPI_F32)
    call("drq_ln_tanh_bwd", ws.dhA.data_ptr(), Fd, ws.hA.data_ptr(), Fd, ws.xhatA.data_ptr(), ws.rstdA.data_ptr(),
         pa("trunk.1.weight"), ws.dz.data_ptr(), ga("trunk.1.weight"), ga("trunk.1.bias"), bw.dz.data_ptr(), ldF,
         B, Fd, s)
    gemm(bw.dz.data_ptr(), ldF, 1, feat_o, REPR_DIM, 1, ga("trunk.0.weight"), REPR_DIM, Fd, REPR_DIM, B,
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


def act_body(agent, w, n, sample):
    """encoder + actor for act() at batch n on the tensor-core path."""
    st, s = agent._bf16, _stream()
    A, Fd, H = agent.action_dim, agent.feature_dim, agent.hidden_dim
    if sample:
        call("drq_rng_normal_f32", agent._seed, agent._counter.data_ptr(), w["eps"].data_ptr(), n * A, s)
        call("drq_counter_advance", agent._counter.data_ptr(), s)
    be = lambda i: agent._p("encoder", f"convnet.{i}.bias")
    acts = [a.data_ptr() for a in w["acts_b"]]
    call("drq_conv1_fwd_bf16", w["obs"].data_ptr(), None, st.conv1_w.data_ptr(), be(0), acts[0], n,
         agent.obs_shape[0], agent.aug.pad, s)
    call("drq_conv3x3_fwd_bf16", acts[0], st.conv_wf[0].data_ptr(), be(2), acts[1], n, 39, 0, s)
    call("drq_conv3x3_fwd_bf16", acts[1], st.conv_wf[1].data_ptr(), be(4), acts[2], n, 37, 0, s)
    call("drq_conv3x3_fwd_bf16", acts[2], st.conv_wf[2].data_ptr(), be(6), w["feat_b"].data_ptr(), n, 35, 1, s)
    pa = lambda k: agent._p("actor", k)
    S = w["S_b"]
    gemm(w["feat_b"].data_ptr(), REPR_DIM, 0, st.actor.ptr("trunk"), REPR_DIM, 0, w["partial_b"].data_ptr(), Fd, n,
         Fd, REPR_DIM, TEPI_F32, splitk=S, bs=(0, 0, n * Fd, 0, 0), bn=64)
    call("drq_ln_tanh_fwd", w["partial_b"].data_ptr(), S, n * Fd, pa("trunk.0.bias"), pa("trunk.1.weight"),
         pa("trunk.1.bias"), w["h"].data_ptr(), Fd, None, None, w["h_b"].data_ptr(), st.ldF, n, Fd, 1e-5, s)
    gemm(w["h_b"].data_ptr(), st.ldF, 0, st.actor.ptr("policy.0"), st.ldF, 0, w["p1_b"].data_ptr(), H, n, H, Fd,
         TEPI_RELU_BF16, bias=pa("policy.0.bias"))
    gemm(w["p1_b"].data_ptr(), H, 0, st.actor.ptr("policy.2"), H, 0, w["p2_b"].data_ptr(), H, n, H, H,
         TEPI_RELU_BF16, bias=pa("policy.2.bias"))
    gemm(w["p2_b"].data_ptr(), H, 0, st.actor.ptr("policy.4"), H, 0, w["mu_pre"].data_ptr(), A, n, A, H, TEPI_F32,
         bias=pa("policy.4.bias"), bn=32)
    call("drq_actor_sample", w["mu_pre"].data_ptr(), w["eps"].data_ptr() if sample else None,
         agent._scal_dev.data_ptr() + F32 * 8, 0.0, w["out"].data_ptr(), A, None, None, None, 0, n, A, s)
