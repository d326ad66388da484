"""CPU check of the index arithmetic of csrc/conv2x2_tc.cu (the 2x2-output-block conv kernels): block positions numbered
across images, parity-plane windows with zero fill, the 16 per-window-element weight operands expanded from the compact
[tap*4 + k/8][n][k%8] layout, accumulator columns (oy, ox, co).  The emulation follows the kernel statement by statement
(same formulas, numpy instead of UMMAs) and is compared with torch's conv2d / conv_transpose2d in float64."""
import numpy as np
import pytest
import torch


def region_wy(r):
    return 1 + (r >> 1) if r < 4 else 1 + ((r - 4) >> 1) if r < 8 else 3 * ((r - 8) >> 1) if r < 12 else 3 * ((r - 12) >> 1)


def region_wx(r):
    return 1 + (r & 1) if r < 4 else 3 * ((r - 4) & 1) if r < 8 else 1 + ((r - 8) & 1) if r < 12 else 3 * ((r - 12) & 1)


def region_n(r):
    return 128 if r < 8 else 64 if r < 12 else 32


def region_col(r):
    return 0 if r < 8 else (region_wy(r) // 3) * 64 if r < 12 else ((region_wy(r) // 3) * 2 + region_wx(r) // 3) * 32


def compact_weights(w, dgrad):
    """pack_conv_w_elem (csrc/pack.cuh): [tap][k][n] with (k, n) = (ci, co) forward, (co, ci) data gradient."""
    out = np.zeros((9, 32, 32))
    for co in range(32):
        for ci in range(32):
            for tap in range(9):
                v = w[co, ci, tap // 3, tap % 3]
                if dgrad:
                    out[tap, co, ci] = v
                else:
                    out[tap, ci, co] = v
    return out


def expand(wc, dgrad):
    """expand_weights: region r -> [k = 32][n = region_n(r)]"""
    regs = []
    for r in range(16):
        wy, wx, nt = region_wy(r), region_wx(r), region_n(r)
        B = np.zeros((32, nt))
        for n in range(nt):
            if r < 8:
                oy, ox = n >> 6, (n >> 5) & 1
            elif r < 12:
                oy, ox = wy // 3, n >> 5
            else:
                oy, ox = wy // 3, wx // 3
            co = n & 31
            dy = oy + 2 - wy if dgrad else wy - oy
            dx = ox + 2 - wx if dgrad else wx - ox
            if 0 <= dy <= 2 and 0 <= dx <= 2:
                B[:, n] = wc[dy * 3 + dx, :, co]
        regs.append(B)
    return regs


def emulate(x, w, h_layer_out, dgrad):
    """x [N][32][h][h] float64: the layer input (forward) or the gradient of its output (data gradient)."""
    N = x.shape[0]
    hin = h_layer_out + 2
    pitch = (hin + 1) // 2
    h_in = h_layer_out if dgrad else hin
    h_out = hin if dgrad else h_layer_out
    nrow = pitch + 1 if dgrad else pitch
    pl4 = pitch * nrow
    total = N * pl4
    tiles = (total + 127) // 128
    regs = expand(compact_weights(w, dgrad), dgrad)
    out = np.zeros((N, 32, h_out, h_out))
    written = np.zeros((N, h_out, h_out), bool)

    def pixel(v, py, px):                      # the loader's decode + zero fill
        if v < 0 or v >= total:
            return np.zeros(32)
        n, q = divmod(v, pl4)
        i, j = divmod(q, pitch)
        y, xx = 2 * i + py, 2 * j + px
        if y >= h_in or xx >= h_in:
            return np.zeros(32)
        return x[n, :, y, xx]

    for t in range(tiles):
        v0 = t * 128 - (pitch + 1 if dgrad else 0)
        n_slots = 128 + pitch + 1
        planes = np.zeros((2, 2, n_slots, 32))
        for s in range(n_slots):
            for py in range(2):
                for px in range(2):
                    planes[py, px, s] = pixel(v0 + s, py, px)
        D = np.zeros((128, 128))
        for r in range(16):
            wy, wx, nt, col = region_wy(r), region_wx(r), region_n(r), region_col(r)
            off = (wy >> 1) * pitch + (wx >> 1)
            A = planes[wy & 1, wx & 1, off:off + 128]
            D[:, col:col + nt] += A @ regs[r]
        for m in range(128):
            v = t * 128 + m
            if v >= total:
                continue
            n, q = divmod(v, pl4)
            i, j = divmod(q, pitch)
            for oy in range(2):
                for ox in range(2):
                    y, xx = 2 * i + oy, 2 * j + ox
                    if y < h_out and xx < h_out:
                        assert not written[n, y, xx]
                        written[n, y, xx] = True
                        out[n, :, y, xx] = D[m, (oy * 2 + ox) * 32:(oy * 2 + ox) * 32 + 32]
    assert written.all()
    return out


@pytest.mark.parametrize("hout,N", [(39, 2), (36, 1), (5, 3)])
def test_conv2x2_forward_indexing(hout, N):
    g = torch.Generator().manual_seed(hout)
    x = torch.rand(N, 32, hout + 2, hout + 2, generator=g, dtype=torch.float64)
    w = torch.rand(32, 32, 3, 3, generator=g, dtype=torch.float64) - 0.5
    want = torch.nn.functional.conv2d(x, w).numpy()
    got = emulate(x.numpy(), w.numpy(), hout, False)
    assert np.abs(got - want).max() <= 1e-10


@pytest.mark.parametrize("hout,N", [(39, 2), (36, 1), (5, 3)])
def test_conv2x2_dgrad_indexing(hout, N):
    g = torch.Generator().manual_seed(100 + hout)
    d = torch.rand(N, 32, hout, hout, generator=g, dtype=torch.float64) - 0.5
    w = torch.rand(32, 32, 3, 3, generator=g, dtype=torch.float64) - 0.5
    want = torch.nn.functional.conv_transpose2d(d, w).numpy()
    got = emulate(d.numpy(), w.numpy(), hout, True)
    assert np.abs(got - want).max() <= 1e-10


def test_magic_division_exact():
    """x / d = umulhi(x, floor(2^32 / d) + 1) for every position the kernels can see (x * d < 2^32)."""
    for d in [19, 20, 21, 361, 380, 400, 420, 441, 462]:
        m = (1 << 32) // d + 1
        xs = np.arange(0, 1 << 23, 7, dtype=np.uint64)
        assert np.array_equal((xs * np.uint64(m)) >> np.uint64(32), xs // np.uint64(d))
        edge = np.array([k * d + e for k in range(1, 20000, 37) for e in (-1, 0, 1)], dtype=np.uint64)
        assert np.array_equal((edge * np.uint64(m)) >> np.uint64(32), edge // np.uint64(d))
