"""CPU oracle for the DrQ-v2 agent-update hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``drqv2_b200/`` imports this module; it is
used by ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs as the checker and the timed CPU baseline.

It is an independent restatement (numpy for the byte/integer work, torch-CPU fp32 or
fp64 for the floating-point work) of the algorithm in the reference
(johannah/drqv2: drqv2.py, utils.py, replay_buffer.py, dmc.py).  Each function cites
the reference lines it follows.

Pinning: the reference has no tests or golden vectors of its own, so
the oracle is pinned against outputs of the reference itself, imported in the build
container by ``tests/golden/make_golden.py`` and committed under ``tests/golden/``
(see ``tests/test_oracle_golden.py``).
"""
from __future__ import annotations

import math
import re
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

IMG = 84
REPR_DIM = 32 * 35 * 35  # drqv2.py:53


# --------------------------------------------------------------------------- replay
def stack_rows(t: int, stack: int = 3):
    """Rows of single frames forming the stacked observation of row ``t`` of an episode.

    dmc.py:98-109 — a deque of the last ``stack`` frames, filled with the reset frame
    at row 0; so row t stacks frames max(t-2,0), max(t-1,0), t.
    """
    return [max(t - (stack - 1 - j), 0) for j in range(stack)]


def nstep_sample(frames, action, reward, discount, idx, nstep, gamma, stack=3):
    """One sample of ReplayBuffer._sample (replay_buffer.py:150-160) from per-row
    single frames.

    frames u8 [T+1, C, 84, 84] (row 0 = reset), action f32 [T+1, A],
    reward/discount f32 [T+1, 1].  Returns (obs, action, reward, discount, next_obs)
    with the reference's dtypes/shapes; reward/discount follow the sequential fp32
    chain of replay_buffer.py:154-159.
    """
    obs = np.concatenate([frames[r] for r in stack_rows(idx - 1, stack)], axis=0)
    nxt = np.concatenate([frames[r] for r in stack_rows(idx + nstep - 1, stack)], axis=0)
    rew = np.zeros_like(reward[idx])
    disc = np.ones_like(discount[idx])
    g32 = np.float32(gamma)
    for i in range(nstep):
        rew = (rew + disc * reward[idx + i]).astype(np.float32)
        disc = (disc * (discount[idx + i] * g32).astype(np.float32)).astype(np.float32)
    return obs, action[idx].copy(), rew, disc, nxt


def ring_gather(ring_frames, ring_action, ring_reward, ring_discount, ep_start, idx, nstep, gamma,
                stack=3):
    """Batched gather from a ring (slots modulo capacity); mirrors drq_ring_gather_nstep."""
    cap = ring_frames.shape[0]
    B = len(idx)
    C = ring_frames.shape[1]
    obs = np.empty((B, C * stack, IMG, IMG), np.uint8)
    nxt = np.empty_like(obs)
    act = np.empty((B, ring_action.shape[1]), np.float32)
    rew = np.empty((B, 1), np.float32)
    disc = np.empty((B, 1), np.float32)
    g32 = np.float32(gamma)
    for b in range(B):
        s0, i0 = int(ep_start[b]), int(idx[b])
        for j, r in enumerate(stack_rows(i0 - 1, stack)):
            obs[b, j * C:(j + 1) * C] = ring_frames[(s0 + r) % cap]
        for j, r in enumerate(stack_rows(i0 + nstep - 1, stack)):
            nxt[b, j * C:(j + 1) * C] = ring_frames[(s0 + r) % cap]
        act[b] = ring_action[(s0 + i0) % cap]
        r_acc, d_acc = np.float32(0.0), np.float32(1.0)
        for i in range(nstep):
            s = (s0 + i0 + i) % cap
            r_acc = np.float32(r_acc + np.float32(d_acc * ring_reward[s]))
            d_acc = np.float32(d_acc * np.float32(ring_discount[s] * g32))
        rew[b, 0], disc[b, 0] = r_acc, d_acc
    return obs, act, rew, disc, nxt


# --------------------------------------------------------------------------- RNG (Philox4x32-10)
_M0, _M1, _W0, _W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85


def philox4x32(seed: int, stream: int, index: np.ndarray) -> np.ndarray:
    """Counter-based generator used by the CUDA sampler/draw kernels (csrc/common.cuh)."""
    index = np.asarray(index, dtype=np.uint64)
    c = [(index & np.uint64(0xFFFFFFFF)), (index >> np.uint64(32)),
         np.full_like(index, stream & 0xFFFFFFFF), np.full_like(index, (stream >> 32) & 0xFFFFFFFF)]
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(_M0) * c[0]
        p1 = np.uint64(_M1) * c[2]
        h0, l0, h1, l1 = p0 >> np.uint64(32), p0 & mask, p1 >> np.uint64(32), p1 & mask
        c = [h1 ^ c[1] ^ np.uint64(k0), l1, h0 ^ c[3] ^ np.uint64(k1), l0]
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return np.stack(c, axis=-1).astype(np.uint64)


def sample_indices(ep_table, nstep, seed, counter, B):
    """Episode ~ U{0..E-1}, idx ~ U{1..len-nstep+1} (replay_buffer.py:96-98,150) with the
    multiply-shift range reduction of drq_ring_sample."""
    ep_table = np.asarray(ep_table, np.int64)
    r = philox4x32(seed, (counter << 3) | 0, np.arange(B))
    e = (r[:, 0] * np.uint64(len(ep_table))) >> np.uint64(32)
    start, length = ep_table[e.astype(np.int64), 0], ep_table[e.astype(np.int64), 1]
    span = (length - nstep + 1).astype(np.uint64)
    idx = ((r[:, 1] * span) >> np.uint64(32)).astype(np.int64) + 1
    return start.astype(np.int32), idx.astype(np.int32)


def update_shifts(seed, counter, pad, B):
    r = philox4x32(seed, (counter << 3) | 1, np.arange(B))
    rng = np.uint64(2 * pad + 1)
    s = ((r * rng) >> np.uint64(32)).astype(np.int32)
    return s[:, 0:2].copy(), s[:, 2:4].copy()


# --------------------------------------------------------------------------- augmentation
def random_shift_exact(x: torch.Tensor, shift: torch.Tensor, pad: int = 4) -> torch.Tensor:
    """RandomShiftsAug as the integer translation it approximates (drqv2.py:19-45):
    out[n,c,r,col] = x[n,c,clamp(r+sy-pad), clamp(col+sx-pad)], shift[n] = (sx, sy)
    — shift[...,0] moves columns, shift[...,1] rows."""
    n, c, h, w = x.shape
    assert h == w  # drqv2.py:21
    sx = shift[:, 0].long().view(n, 1)
    sy = shift[:, 1].long().view(n, 1)
    rows = (torch.arange(h).view(1, h) + sy - pad).clamp(0, h - 1)  # [n,h]
    cols = (torch.arange(w).view(1, w) + sx - pad).clamp(0, w - 1)  # [n,w]
    idx_n = torch.arange(n).view(n, 1, 1, 1)
    idx_c = torch.arange(c).view(1, c, 1, 1)
    return x[idx_n, idx_c, rows.view(n, 1, h, 1), cols.view(n, 1, 1, w)]


def random_shift_grid_sample(x: torch.Tensor, shift: torch.Tensor, pad: int = 4) -> torch.Tensor:
    """The reference's float implementation (drqv2.py:19-45), with the randint draw
    injected: replicate-pad, base grid, shift*2/(h+2pad), bilinear grid_sample."""
    n, c, h, w = x.shape
    assert h == w
    xp = F.pad(x, (pad,) * 4, "replicate")
    eps = 1.0 / (h + 2 * pad)
    arange = torch.linspace(-1.0 + eps, 1.0 - eps, h + 2 * pad, dtype=x.dtype)[:h]
    arange = arange.unsqueeze(0).repeat(h, 1).unsqueeze(2)
    base = torch.cat([arange, arange.transpose(1, 0)], dim=2).unsqueeze(0).repeat(n, 1, 1, 1)
    s = shift.to(x.dtype).view(n, 1, 1, 2) * (2.0 / (h + 2 * pad))
    return F.grid_sample(xp, base + s, padding_mode="zeros", align_corners=False)


# --------------------------------------------------------------------------- schedule
def schedule(schdl, step):
    """utils.py:129-149."""
    try:
        return float(schdl)
    except ValueError:
        m = re.match(r"linear\((.+),(.+),(.+)\)", schdl)
        if m:
            init, final, duration = [float(g) for g in m.groups()]
            mix = np.clip(step / duration, 0.0, 1.0)
            return (1.0 - mix) * init + mix * final
        m = re.match(r"step_linear\((.+),(.+),(.+),(.+),(.+)\)", schdl)
        if m:
            init, final1, duration1, final2, duration2 = [float(g) for g in m.groups()]
            if step <= duration1:
                mix = np.clip(step / duration1, 0.0, 1.0)
                return (1.0 - mix) * init + mix * final1
            mix = np.clip((step - duration1) / duration2, 0.0, 1.0)
            return (1.0 - mix) * final1 + mix * final2
    raise NotImplementedError(schdl)


# --------------------------------------------------------------------------- parameters
def param_shapes(cin, A, Fdim, H):
    """Names/shapes in the order of the reference modules' .parameters()
    (drqv2.py:55-59, 74-81, 100-111)."""
    enc = OrderedDict()
    for i, ci in zip((0, 2, 4, 6), (cin, 32, 32, 32)):
        enc[f"convnet.{i}.weight"] = (32, ci, 3, 3)
        enc[f"convnet.{i}.bias"] = (32,)
    actor = OrderedDict([
        ("trunk.0.weight", (Fdim, REPR_DIM)), ("trunk.0.bias", (Fdim,)),
        ("trunk.1.weight", (Fdim,)), ("trunk.1.bias", (Fdim,)),
        ("policy.0.weight", (H, Fdim)), ("policy.0.bias", (H,)),
        ("policy.2.weight", (H, H)), ("policy.2.bias", (H,)),
        ("policy.4.weight", (A, H)), ("policy.4.bias", (A,)),
    ])
    critic = OrderedDict([
        ("trunk.0.weight", (Fdim, REPR_DIM)), ("trunk.0.bias", (Fdim,)),
        ("trunk.1.weight", (Fdim,)), ("trunk.1.bias", (Fdim,)),
    ])
    for q in ("Q1", "Q2"):
        critic[f"{q}.0.weight"] = (H, Fdim + A)
        critic[f"{q}.0.bias"] = (H,)
        critic[f"{q}.2.weight"] = (H, H)
        critic[f"{q}.2.bias"] = (H,)
        critic[f"{q}.4.weight"] = (1, H)
        critic[f"{q}.4.bias"] = (1,)
    return OrderedDict(encoder=enc, actor=actor, critic=critic)


def synthetic_params(cin, A, Fdim, H, seed=0, dtype=torch.float32):
    """Deterministic, machine-independent parameters (torch CPU generator): scaled
    uniform weights of roughly the magnitude of the reference's orthogonal init
    (utils.py:52-61), non-zero biases so that every bias path is exercised."""
    g = torch.Generator().manual_seed(seed)
    out = OrderedDict()
    for net, shapes in param_shapes(cin, A, Fdim, H).items():
        d = OrderedDict()
        for name, shp in shapes.items():
            if name.endswith("weight") and len(shp) > 1:
                fan_in = int(np.prod(shp[1:]))
                fan_out = shp[0]
                gain = math.sqrt(2.0) if len(shp) == 4 else 1.0
                scale = gain * math.sqrt(3.0 / max(fan_in, fan_out))  # ~ orthogonal-init row norm
                t = (torch.rand(shp, generator=g, dtype=torch.float32) * 2 - 1) * scale
            elif name.startswith("trunk.1.weight"):
                t = 1.0 + 0.1 * (torch.rand(shp, generator=g) * 2 - 1)  # LayerNorm gamma
            else:
                t = 0.05 * (torch.rand(shp, generator=g) * 2 - 1)
            d[name] = t.to(dtype)
        out[net] = d
    out["critic_target"] = OrderedDict((k, v.clone()) for k, v in out["critic"].items())  # drqv2.py:145
    return out


# --------------------------------------------------------------------------- networks (functional)
def encoder_fwd(p, obs):
    """Encoder.forward drqv2.py:63-67.  obs float [N,C,84,84] on the 0..255 scale."""
    h = obs / 255.0 - 0.5
    for i, stride in zip((0, 2, 4, 6), (2, 1, 1, 1)):
        h = F.relu(F.conv2d(h, p[f"convnet.{i}.weight"], p[f"convnet.{i}.bias"], stride=stride))
    return h.reshape(h.shape[0], -1)


def trunk_fwd(p, feat):
    """Linear + LayerNorm + Tanh, drqv2.py:74-75,100-101."""
    z = F.linear(feat, p["trunk.0.weight"], p["trunk.0.bias"])
    z = F.layer_norm(z, (z.shape[-1],), p["trunk.1.weight"], p["trunk.1.bias"], eps=1e-5)
    return torch.tanh(z)


def actor_mu(p, feat):
    """Actor.forward up to mu, drqv2.py:85-89."""
    h = trunk_fwd(p, feat)
    h = F.relu(F.linear(h, p["policy.0.weight"], p["policy.0.bias"]))
    h = F.relu(F.linear(h, p["policy.2.weight"], p["policy.2.bias"]))
    return torch.tanh(F.linear(h, p["policy.4.weight"], p["policy.4.bias"]))


def critic_q(p, feat, action):
    """Critic.forward drqv2.py:115-121."""
    h = trunk_fwd(p, feat)
    x = torch.cat([h, action], dim=-1)
    qs = []
    for q in ("Q1", "Q2"):
        y = F.relu(F.linear(x, p[f"{q}.0.weight"], p[f"{q}.0.bias"]))
        y = F.relu(F.linear(y, p[f"{q}.2.weight"], p[f"{q}.2.bias"]))
        qs.append(F.linear(y, p[f"{q}.4.weight"], p[f"{q}.4.bias"]))
    return qs[0], qs[1]


def truncated_normal_sample(mu, std, eps, clip):
    """utils.TruncatedNormal.sample utils.py:117-126 with the N(0,1) draw injected;
    value = clamp(mu + clip(eps*std)), gradient w.r.t. mu = 1 (straight-through)."""
    e = eps * (torch.ones_like(mu) * std)
    if clip is not None:
        e = torch.clamp(e, -clip, clip)
    x = mu + e
    clamped = torch.clamp(x, -1.0 + 1e-6, 1.0 - 1e-6)
    return x - x.detach() + clamped.detach()


# --------------------------------------------------------------------------- bf16-faithful variants
# The tensor-core mode of the CUDA path (drqv2_b200/_bf16.py) keeps fp32 master parameters and fp32
# accumulation but STORES its GEMM / conv operands in bf16: the weights' operand copies, every hidden
# activation, the features, and the gradient operands of the backward GEMMs.  These variants restate the
# same reference functions (file:line cited per function) with a round-to-bf16 at exactly those points, so
# that a comparison with the kernels is not dominated by ReLU units that sit on the other side of zero
# once an activation is rounded (tests/test_gpu_bf16.py).  Math between the rounding points is in self.dtype
# (float64 in the tests).
class _RoundBf16(torch.autograd.Function):
    """value -> nearest bf16, gradient passed through (an operand copy of an fp32 master value)"""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g


class _RoundGradBf16(torch.autograd.Function):
    """identity whose incoming gradient is rounded to bf16 (a gradient stored as a bf16 GEMM operand)"""

    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(g.dtype)


def rb(x):
    return _RoundBf16.apply(x)


def gb(x):
    return _RoundGradBf16.apply(x)


def encoder_fwd_bf16(p, obs):
    """Encoder.forward drqv2.py:63-67 as the tensor-core path computes it: exact pixels (conv1 feeds the
    integers x - 128 and folds the affine map into its epilogue, csrc/conv1_tc.cu), bf16 weights, every
    post-ReLU activation stored in bf16, every pre-activation gradient stored in bf16."""
    h = obs / 255.0 - 0.5
    for i, stride in zip((0, 2, 4, 6), (2, 1, 1, 1)):
        pre = gb(F.conv2d(h, rb(p[f"convnet.{i}.weight"]), p[f"convnet.{i}.bias"], stride=stride))
        h = rb(F.relu(pre))
    return h.reshape(h.shape[0], -1)


def trunk_fwd_bf16(p, feat):
    """drqv2.py:74-75,100-101: bf16 features x bf16 weight, fp32 bias / LayerNorm / tanh; the gradient of the
    Linear output feeds the weight- and data-gradient GEMMs in bf16, the bias gradient in fp32."""
    z = gb(F.linear(feat, rb(p["trunk.0.weight"]))) + p["trunk.0.bias"]
    z = F.layer_norm(z, (z.shape[-1],), p["trunk.1.weight"], p["trunk.1.bias"], eps=1e-5)
    return torch.tanh(z)


def _hidden_bf16(x, w, b):
    return rb(F.relu(gb(F.linear(x, rb(w), b))))


def actor_mu_bf16(p, feat):
    """Actor.forward drqv2.py:85-89."""
    h = rb(trunk_fwd_bf16(p, feat))
    h = _hidden_bf16(h, p["policy.0.weight"], p["policy.0.bias"])
    h = _hidden_bf16(h, p["policy.2.weight"], p["policy.2.bias"])
    return torch.tanh(gb(F.linear(h, rb(p["policy.4.weight"]))) + p["policy.4.bias"])


def critic_q_bf16(p, feat, action):
    """Critic.forward drqv2.py:115-121; the scalar heads read the fp32 weights (csrc/heads.cu q_head_*)."""
    h = trunk_fwd_bf16(p, feat)
    x = rb(torch.cat([h, action], dim=-1))
    qs = []
    for q in ("Q1", "Q2"):
        y = _hidden_bf16(x, p[f"{q}.0.weight"], p[f"{q}.0.bias"])
        y = _hidden_bf16(y, p[f"{q}.2.weight"], p[f"{q}.2.bias"])
        qs.append(F.linear(y, p[f"{q}.4.weight"], p[f"{q}.4.bias"]))
    return qs[0], qs[1]


# --------------------------------------------------------------------------- optimiser
def adam_scalars(lr, t, beta1=0.9, beta2=0.999, eps=1e-8):
    """Host-side float64 scalar math of torch/optim/adam.py:531-547, in the order of the
    device `scalars` array of drq_adam_step."""
    bc1 = 1 - beta1 ** t
    bc2 = 1 - beta2 ** t
    return np.array([1 - beta1, beta2, 1 - beta2, bc2 ** 0.5, eps, -(lr / bc1), 0, 0], dtype=np.float32)


def adam_step(p, g, m, v, lr, t, beta1=0.9, beta2=0.999, eps=1e-8):
    """torch.optim.Adam single-tensor math (adam.py:457,476,531-547); in place."""
    m.lerp_(g, 1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    bc1 = 1 - beta1 ** t
    bc2 = 1 - beta2 ** t
    denom = (v.sqrt() / (bc2 ** 0.5)).add_(eps)
    p.addcdiv_(m, denom, value=-(lr / bc1))


def soft_update(p, tp, tau):
    """utils.soft_update_params utils.py:42-45."""
    tp.copy_(tau * p + (1 - tau) * tp)


# --------------------------------------------------------------------------- one update
class OracleAgent:
    """Functional restatement of DrQV2Agent (drqv2.py:124-262) over explicit parameter
    dicts, with every random draw injected.  dtype float32 = REF-X (exact shift),
    float64 = REF-64 of SURVEY §8c."""

    def __init__(self, params, lr, critic_target_tau, stddev_schedule, stddev_clip,
                 dtype=torch.float32, aug="exact", operands="exact"):
        """operands="bf16": the bf16-faithful variants above (rounding where the tensor-core path stores
        bf16 operands); "exact": the reference's arithmetic in `dtype`."""
        self.dtype = dtype
        assert operands in ("exact", "bf16")
        self.operands = operands
        bf = operands == "bf16"
        self._encoder_fwd = encoder_fwd_bf16 if bf else encoder_fwd
        self._actor_mu = actor_mu_bf16 if bf else actor_mu
        self._critic_q = critic_q_bf16 if bf else critic_q
        self.p = OrderedDict((net, OrderedDict((k, v.detach().clone().to(dtype)) for k, v in d.items()))
                             for net, d in params.items())
        self.lr, self.tau = lr, critic_target_tau
        self.stddev_schedule, self.stddev_clip = stddev_schedule, stddev_clip
        self.aug = aug
        self.m = {net: OrderedDict((k, torch.zeros_like(v)) for k, v in self.p[net].items())
                  for net in ("encoder", "actor", "critic")}
        self.v = {net: OrderedDict((k, torch.zeros_like(v)) for k, v in self.p[net].items())
                  for net in ("encoder", "actor", "critic")}
        self.t = {"encoder": 0, "actor": 0, "critic": 0}
        self.grads = {}

    def _aug(self, x, shift):
        if self.aug == "exact":
            return random_shift_exact(x, shift)
        return random_shift_grid_sample(x, shift)

    def _opt_step(self, net, grads):
        self.t[net] += 1
        for k, g in grads.items():
            adam_step(self.p[net][k], g, self.m[net][k], self.v[net][k], self.lr, self.t[net])

    def act(self, obs_u8, step, eval_mode, eps=None):
        """drqv2.py:164-175 (without the uniform-exploration overwrite)."""
        with torch.no_grad():
            x = torch.as_tensor(obs_u8).to(self.dtype)
            if x.dim() == 3:
                x = x.unsqueeze(0)
            feat = self._encoder_fwd(self.p["encoder"], x)
            mu = self._actor_mu(self.p["actor"], feat)
            if eval_mode:
                return mu
            std = schedule(self.stddev_schedule, step)
            return truncated_normal_sample(mu, std, eps.to(self.dtype), None)

    def update_critic(self, feat, action, reward, discount, feat_next, step, eps_critic, enc=None):
        """drqv2.py:177-204 on encoded features.  `enc`: the encoder parameters `feat` is attached to (as inside
        update(), drqv2.py:244) - they receive the critic loss' gradient and encoder_opt steps (drqv2.py:202);
        with detached features (enc=None) only critic_opt steps."""
        dt = self.dtype
        metrics = {}
        action, reward, discount = action.to(dt), reward.to(dt), discount.to(dt)
        cri = {k: v.requires_grad_(True) for k, v in self.p["critic"].items()}
        act, tgt = self.p["actor"], self.p["critic_target"]
        std = schedule(self.stddev_schedule, step)
        with torch.no_grad():
            mu_n = self._actor_mu(act, feat_next)
            next_action = truncated_normal_sample(mu_n, std, eps_critic.to(dt), self.stddev_clip)
            tq1, tq2 = self._critic_q(tgt, feat_next, next_action)
            target_q = reward + discount * torch.min(tq1, tq2)
        q1, q2 = self._critic_q(cri, feat, action)
        critic_loss = F.mse_loss(q1, target_q) + F.mse_loss(q2, target_q)
        metrics.update(critic_target_q=target_q.mean().item(), critic_q1=q1.mean().item(),
                       critic_q2=q2.mean().item(), critic_loss=critic_loss.item())
        self.stage = dict(feat=feat.detach().clone(), feat_next=feat_next.clone(),
                          next_action=next_action.clone(), target_q=target_q.clone(),
                          q1=q1.detach().clone(), q2=q2.detach().clone())
        enc = enc or {}
        gs = torch.autograd.grad(critic_loss, list(enc.values()) + list(cri.values()))
        g_enc = OrderedDict(zip(list(enc), gs[:len(enc)]))
        g_cri = OrderedDict(zip(list(cri), gs[len(enc):]))
        for d in (enc, cri):
            for v in d.values():
                v.requires_grad_(False)
        self._opt_step("critic", g_cri)
        if enc:
            self._opt_step("encoder", g_enc)
        self.grads.update(encoder=g_enc, critic=g_cri)
        return metrics

    def update_actor(self, feat, step, eps_actor):
        """drqv2.py:206-228 on detached features, with the current (stepped) critic."""
        dt = self.dtype
        act = {k: v.requires_grad_(True) for k, v in self.p["actor"].items()}
        std = schedule(self.stddev_schedule, step)
        feat_d = feat.detach()
        mu = self._actor_mu(act, feat_d)
        a = truncated_normal_sample(mu, std, eps_actor.to(dt), self.stddev_clip)
        var = torch.as_tensor(std, dtype=dt) ** 2
        log_prob = (-((a - mu) ** 2) / (2 * var) - math.log(std) - math.log(math.sqrt(2 * math.pi)))
        log_prob = log_prob.sum(-1, keepdim=True)
        aq1, aq2 = self._critic_q(self.p["critic"], feat_d, a)
        actor_loss = -torch.min(aq1, aq2).mean()
        ga = torch.autograd.grad(actor_loss, list(act.values()))
        g_act = OrderedDict(zip(list(act), ga))
        for v in act.values():
            v.requires_grad_(False)
        self.stage.update(actor_action=a.detach().clone(), actor_q1=aq1.detach().clone(),
                          actor_q2=aq2.detach().clone())
        self._opt_step("actor", g_act)
        self.grads.update(actor=g_act)
        return dict(actor_loss=actor_loss.item(), actor_logprob=log_prob.mean().item(),
                    actor_ent=float(mu.shape[-1] * (0.5 + 0.5 * math.log(2 * math.pi) + math.log(std))))

    def encode(self, obs_u8, shift, enc=None):
        """aug + encoder (drqv2.py:241-246) with the shift draw injected"""
        x = self._aug(torch.as_tensor(obs_u8).to(self.dtype), shift)
        return self._encoder_fwd(enc if enc is not None else self.p["encoder"], x)

    def update(self, obs_u8, action, reward, discount, next_obs_u8, step, shift_obs, shift_next,
               eps_critic, eps_actor):
        """One DrQV2Agent.update (drqv2.py:230-262) on an explicit batch.  Returns the
        metrics dict; gradients of the three nets are kept in self.grads."""
        dt = self.dtype
        metrics = {}
        enc = {k: v.requires_grad_(True) for k, v in self.p["encoder"].items()}
        # augment + encode, drqv2.py:241-246
        feat = self.encode(obs_u8, shift_obs, enc)
        with torch.no_grad():
            feat_next = self.encode(next_obs_u8, shift_next, enc)
        metrics["batch_reward"] = reward.to(dt).mean().item()
        # update_critic, drqv2.py:249-250; update_actor on detached features, drqv2.py:256
        metrics.update(self.update_critic(feat, action, reward, discount, feat_next, step, eps_critic, enc=enc))
        metrics.update(self.update_actor(feat.detach(), step, eps_actor))
        # soft target update, drqv2.py:259-260
        with torch.no_grad():
            for k in self.p["critic"]:
                soft_update(self.p["critic"][k], self.p["critic_target"][k], self.tau)
        return metrics


def synthetic_batch(B, A, cin=9, seed=1):
    """Synthetic replay batch of SURVEY §8d: uniform u8 pixels, action~U(-1,1),
    reward~U(0,1), discount = 0.99^3 (fp32 chain value), plus the four injected draws."""
    g = torch.Generator().manual_seed(seed)
    obs = torch.randint(0, 256, (B, cin, IMG, IMG), dtype=torch.uint8, generator=g)
    nxt = torch.randint(0, 256, (B, cin, IMG, IMG), dtype=torch.uint8, generator=g)
    action = torch.rand(B, A, generator=g) * 2 - 1
    reward = torch.rand(B, 1, generator=g)
    d = np.float32(1.0)
    for _ in range(3):
        d = np.float32(d * np.float32(np.float32(1.0) * np.float32(0.99)))
    discount = torch.full((B, 1), float(d))
    shift_obs = torch.randint(0, 9, (B, 2), generator=g, dtype=torch.int32)
    shift_next = torch.randint(0, 9, (B, 2), generator=g, dtype=torch.int32)
    eps_c = torch.randn(B, A, generator=g)
    eps_a = torch.randn(B, A, generator=g)
    return dict(obs=obs, action=action, reward=reward, discount=discount, next_obs=nxt,
                shift_obs=shift_obs, shift_next=shift_next, eps_critic=eps_c, eps_actor=eps_a)
