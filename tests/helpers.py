"""Shared test helpers: the integer formulas the golden inputs were generated with
(tests/golden/make_golden.py) and probe positions."""
import numpy as np


def frame_formula(e, t):
    c = np.arange(3).reshape(3, 1, 1)
    y = np.arange(84).reshape(1, 84, 1)
    x = np.arange(84).reshape(1, 1, 84)
    return ((31 * e + 17 * t + 7 * c + 3 * y + 5 * x + (x * y) % 11 + (e + 1) * (t + 2) * (x + y) % 13) % 256).astype(np.uint8)


def scalar_formula(e, t, A):
    action = (((np.arange(A) * 37 + e * 11 + t * 5) % 200) / 100.0 - 1.0).astype(np.float32)
    reward = np.float32(((e * 7 + t * 13) % 97) / 97.0)
    discount = np.float32(1.0 if (t + e) % 5 else 0.9)
    return action, reward, discount


def episode_arrays(e, T, A):
    """Per-row single frames + scalars of golden episode e (rows 0..T; row 0 = reset dummy:
    action 0, reward 0, discount 1 — dmc.py:160-168)."""
    frames = np.stack([frame_formula(e, t) for t in range(T + 1)])
    action = np.zeros((T + 1, A), np.float32)
    reward = np.zeros((T + 1, 1), np.float32)
    discount = np.ones((T + 1, 1), np.float32)
    for t in range(1, T + 1):
        a, r, d = scalar_formula(e, t, A)
        action[t], reward[t, 0], discount[t, 0] = a, r, d
    return frames, action, reward, discount


def aug_input(N, C):
    n = np.arange(N).reshape(N, 1, 1, 1)
    c = np.arange(C).reshape(1, C, 1, 1)
    y = np.arange(84).reshape(1, 1, 84, 1)
    x = np.arange(84).reshape(1, 1, 1, 84)
    return ((n * 53 + c * 29 + y * y * 3 + x * 7 + (x * y) % 17) % 256).astype(np.uint8)


def rel_l2(a, b):
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


# ---- bf16 "WB" layout helpers (see include/drqv2_b200.h): torch-side conversions for tests
PLB, GUARD, WB_SLACK = 1776, 88, 128


def wb_from_nchw(x):
    """x float [N,32,H,W] (H,W <= 41) -> WB bf16 tensor [4][N*PLB+SLACK][8] on x.device."""
    import torch
    N, C, H, W = x.shape
    wide = torch.zeros(N, C, 41, 41, dtype=torch.float32, device=x.device)
    wide[:, :, :H, :W] = x
    rows = torch.zeros(N, PLB, C, dtype=torch.float32, device=x.device)
    rows[:, GUARD:GUARD + 1681, :] = wide.view(N, C, 1681).permute(0, 2, 1)
    wb = torch.zeros(4, N * PLB + WB_SLACK, 8, dtype=torch.bfloat16, device=x.device)
    wb[:, :N * PLB, :] = rows.view(N * PLB, 4, 8).permute(1, 0, 2).to(torch.bfloat16)
    return wb.contiguous()


def nchw_from_wb(wb, N, H, W):
    """inverse of wb_from_nchw: float32 [N,32,H,W] (only the valid region)."""
    rows = wb[:, :N * PLB, :].float().permute(1, 0, 2).reshape(N, PLB, 32)
    wide = rows[:, GUARD:GUARD + 1681, :].permute(0, 2, 1).reshape(N, 32, 41, 41)
    return wide[:, :, :H, :W].contiguous()
