"""CPU emulation of the row producers of csrc/conv1_tc.cu (conv1_planes_kernel, conv1_tc_kernel): the software pipeline that
requests an image's shift / ring index / episode start four tiles ahead must hand every tile the values of ITS image, and
the ring slot computed by compare-and-subtract in 32 bits must be the reference's modulo
(replay_buffer.py:150-153 rows idx - 1 / idx + nstep - 1 of an episode stored from `ep_start`, dmc.py:98-109 frame stack)."""
import numpy as np
import pytest

AHEAD = 4
TILES_PER_IMAGE = 14


def producer_sequence(cta, grid, total_tiles, shift, ring_idx, ep_start, B, nstep, stack, capacity, lane):
    """Restatement of the producer loop of one CTA / lane (frame `lane` of the stack): yields (tile, sy, ring slot)."""
    def fetch(t):
        if t >= total_tiles:
            return (None, None, None)
        n = t // TILES_PER_IMAGE
        b = n - B if n >= B else n
        return (int(shift[n][1]), int(ring_idx[b]), int(ep_start[b]))

    q = [fetch(cta + d * grid) for d in range(AHEAD)]
    t0 = cta
    while t0 < total_tiles:
        for d in range(AHEAD):
            t = t0 + d * grid
            if t >= total_tiles:
                break
            sy, idx, ep = q[d]
            n = t // TILES_PER_IMAGE
            tt = idx + nstep - 1 if n >= B else idx - 1
            r = max(tt - (stack - 1 - lane), 0)
            slot = ep + r
            assert slot < 2 ** 31
            while slot >= capacity:
                slot -= capacity
            q[d] = fetch(t + AHEAD * grid)          # consumed four tiles later, in the same slot
            yield t, sy, slot
        t0 += AHEAD * grid


@pytest.mark.parametrize("grid,images", [(148, 512), (140, 96), (7, 3), (148, 48)])
def test_producer_prefetch_hands_every_tile_its_own_image(grid, images):
    g = np.random.default_rng(images)
    B, nstep, stack, capacity = images // 2 if images % 2 == 0 else images, 3, 3, 1000
    n_img = 2 * B if images % 2 == 0 else B
    total = n_img * TILES_PER_IMAGE
    shift = g.integers(0, 9, (n_img, 2))
    ep_start = g.integers(0, capacity, B)                # episodes may start anywhere in the ring ...
    ring_idx = g.integers(1, 400, B)                     # ... and run past its end (wrap)
    ep_start[0] = capacity - 1
    seen = set()
    for cta in range(min(grid, total)):
        for lane in range(stack):
            for t, sy, slot in producer_sequence(cta, grid, total, shift, ring_idx, ep_start, B, nstep, stack, capacity, lane):
                n = t // TILES_PER_IMAGE
                b = n - B if n >= B else n
                assert sy == shift[n][1]
                row = ring_idx[b] + nstep - 1 if n >= B else ring_idx[b] - 1
                want = (int(ep_start[b]) + max(int(row) - (stack - 1 - lane), 0)) % capacity
                assert slot == want, (t, lane)
                if lane == 0:
                    assert t not in seen
                    seen.add(t)
    assert seen == set(range(total)), "every tile is produced exactly once"
    assert any(int(ep_start[b]) + int(ring_idx[b]) + nstep - 1 >= capacity for b in range(B)), "the case must wrap"
