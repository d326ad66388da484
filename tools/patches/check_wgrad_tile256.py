"""Index-math check of wgrad_tile256_no_gain.patch (no GPU): every gradient position is covered exactly once, the
d rows a tile reads stay in front of the next image's first real row, and no copy leaves the WB buffer."""
PLB, GUARD, SLACK, PW, HALO, TM = 1776, 88, 128, 41, 84, 256
for N in (1, 4, 256):
    for hout in (35, 37, 39):
        n_pos = hout * PW
        ntiles = (n_pos + TM - 1) // TM
        rows_total = N * PLB + SLACK
        covered = [0] * n_pos
        for n in sorted({0, N - 1}):
            for t in range(ntiles):
                p0 = t * TM
                drows = min(TM // 16, (n_pos - p0 + 15) // 16) * 16
                xrows = drows + HALO
                row0 = n * PLB + GUARD + p0
                assert row0 + 2 + xrows <= rows_total and row0 + drows <= rows_total, (N, hout, t)
                assert GUARD + p0 + drows <= PLB + GUARD, (hout, t)
                assert drows + 81 < xrows                  # last K step, dy = 2: row (drows - 1) + 82 of a shifted plane
                if n == 0:
                    for p in range(p0, min(p0 + drows, n_pos)):
                        covered[p] += 1
        assert all(c == 1 for c in covered), hout
print("tile256 index math ok")
