"""CPU check of the index arithmetic of csrc/conv4x1_tc.cu (the conv kernels whose accumulator row is a column of four
output pixels): per-image tiles of three block rows, the TMA box (row blocks x row planes, out-of-range row blocks
zero-filled), window element (wy, wx) -> plane / row offset / column offset, the 18 B operands expanded from the compact
[tap*4 + k/8][n][k%8] layout and the accumulator columns (oy, co).  The emulation follows the kernel statement by
statement (same formulas, numpy instead of TMA + UMMAs).  Everything the kernel may read but must not depend on (columns
beyond the valid width in the forward, image rows 41..43, shared-memory slots past a region) is NaN here."""
import numpy as np
import pytest
import torch

PW = 41


def win_cnt(wy):
    return 1 if wy in (0, 5) else 2 if wy in (1, 4) else 3


def win_oy0(wy):
    return 0 if wy < 2 else wy - 2


def compact_weights(w, dgrad):
    """pack_conv_w_elem (csrc/pack.cuh): [tap][k][n] with (k, n) = (ci, co) forward, (co, ci) data gradient."""
    out = np.zeros((9, 32, 32))
    for tap in range(9):
        m = w[:, :, tap // 3, tap % 3]           # [co][ci]
        out[tap] = m if dgrad else m.T
    return out


def expand(wc, dgrad):
    """expand_weights: (wy, wx) -> [k = 32][n = 32 * cnt(wy)]"""
    regs = {}
    for wx in range(3):
        for wy in range(6):
            cnt = win_cnt(wy)
            B = np.zeros((32, 32 * cnt))
            for n in range(32 * cnt):
                oy, co = win_oy0(wy) + (n >> 5), n & 31
                dy = oy + 2 - wy if dgrad else wy - oy
                dx = 2 - wx if dgrad else wx
                assert 0 <= dy <= 2
                B[:, n] = wc[dy * 3 + dx, :, co]
            regs[wy, wx] = B
    return regs


def emulate(wide, w, h_layer_out, dgrad):
    """wide [N][32][41][41] float64: the WB content of the input buffer (wide plane), NaN where the kernel must not look."""
    N = wide.shape[0]
    h_out = h_layer_out + 2 if dgrad else h_layer_out
    block_rows = (h_out + 3) // 4
    tiles_per_image = (block_rows + 2) // 3
    regs = expand(compact_weights(w, dgrad), dgrad)
    out = np.zeros((N, 32, h_out, h_out))
    written = np.zeros((N, h_out, h_out), bool)
    for t in range(N * tiles_per_image):
        n, tau = divmod(t, tiles_per_image)
        i0 = 3 * tau
        rb0 = i0 - (1 if dgrad else 0)
        # the box: [plane][row block 0..3][x] (+ NaN slots standing for whatever follows the region)
        planes = np.full((4, 4 * PW + 8, 32), np.nan)
        for p in range(4):
            for rr in range(4):
                rb = rb0 + rr
                row = 4 * rb + p
                if rb < 0 or rb >= 11:
                    planes[p, rr * PW:(rr + 1) * PW] = 0.0               # TMA out-of-bounds fill
                elif row < PW:
                    planes[p, rr * PW:(rr + 1) * PW] = wide[n, :, row, :].T
                # rows 41..43: the next image's guard rows -> left NaN
        D = np.zeros((128, 128))
        for wy in (2, 5, 3, 1, 4, 0):
            plane = (wy + 2) & 3 if dgrad else wy & 3
            down = int(wy >= 2) if dgrad else wy >> 2
            cnt, col = win_cnt(wy), 32 * win_oy0(wy)
            for wx in range(3):
                off = down * PW + wx - (2 if dgrad else 0)
                A = np.full((123, 32), np.nan)
                for m in range(123):
                    s = m + off
                    A[m] = planes[plane, s] if s >= 0 else (planes[plane - 1, 4 * PW + s] if plane > 0 else np.nan)
                prod = A @ regs[wy, wx]
                if wx == 0 and wy in (2, 5):
                    D[:123, col:col + 32 * cnt] = prod
                else:
                    D[:123, col:col + 32 * cnt] += prod
        for m in range(123):
            r, x = divmod(m, PW)
            for oy in range(4):
                y = 4 * (i0 + r) + oy
                if y < h_out and x < h_out:
                    assert not written[n, y, x]
                    written[n, y, x] = True
                    out[n, :, y, x] = D[m, oy * 32:oy * 32 + 32]
    assert written.all()
    return out


@pytest.mark.parametrize("hout,N", [(39, 2), (37, 1), (35, 2), (6, 2)])
def test_conv4x1_forward_indexing(hout, N):
    g = torch.Generator().manual_seed(hout)
    hin = hout + 2
    x = torch.rand(N, 32, hin, hin, generator=g, dtype=torch.float64)
    w = torch.rand(32, 32, 3, 3, generator=g, dtype=torch.float64) - 0.5
    want = torch.nn.functional.conv2d(x, w).numpy()
    wide = np.full((N, 32, PW, PW), np.nan)
    wide[:, :, :hin, :hin] = x.numpy()
    got = emulate(wide, w.numpy(), hout, False)
    assert np.isfinite(got).all()
    assert np.abs(got - want).max() <= 1e-10


@pytest.mark.parametrize("hout,N", [(39, 2), (37, 1), (35, 2), (6, 2)])
def test_conv4x1_dgrad_indexing(hout, N):
    g = torch.Generator().manual_seed(100 + hout)
    d = torch.rand(N, 32, hout, hout, generator=g, dtype=torch.float64) - 0.5
    w = torch.rand(32, 32, 3, 3, generator=g, dtype=torch.float64) - 0.5
    want = torch.nn.functional.conv_transpose2d(d, w).numpy()
    wide = np.zeros((N, 32, PW, PW))                 # gradient buffers are zero outside the valid region (conv_tc.cu invariant)
    wide[:, :, :hout, :hout] = d.numpy()
    got = emulate(wide, w.numpy(), hout, True)
    assert np.isfinite(got).all()
    assert np.abs(got - want).max() <= 1e-10


def test_expanded_weight_units_cover_the_buffer_once():
    """the unit decode of expand_weights: u -> (wx, wy, k unit, n) is a bijection onto [0, 4608)"""
    prefix = {0: 0, 1: 1, 2: 3, 3: 6, 4: 9, 5: 11}
    seen = set()
    for u in range(3 * 12 * 128):
        wx, rem = divmod(u, 1536)
        s = rem >> 7
        wy = 0 if s < 1 else 1 if s < 3 else 2 if s < 6 else 3 if s < 9 else 4 if s < 11 else 5
        n32 = 32 * win_cnt(wy)
        r2 = rem - 128 * prefix[wy]
        assert 0 <= r2 < 4 * n32
        ku, n = divmod(r2, n32)
        seen.add((wx, wy, ku, n))
        assert wx * 1536 + 128 * prefix[wy] + ku * n32 + n == u
    assert len(seen) == 3 * 12 * 128
