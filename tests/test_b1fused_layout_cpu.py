"""Layout invariants of the one-kernel SNGAN block 1 (csrc/conv_b1fused.cu), restated in Python: the lane -> pixel maps must
tile every T cell of a unit exactly once, the packed per-thread words of the CH = 64 path must fit their bit fields, the patch
offsets must stay inside the patch buffer, and the CH = 64 gathers must be free of shared-memory bank conflicts for the 8-cell
lane rows (the property profiles/r4_b1fused64.md measures on the GPU).  The GPU tests (tests/test_gpu_b1fused.py) check the
kernel's results bit for bit; this file pins the arithmetic those results rest on without a GPU."""
import itertools

BF_ROW, BF_K8 = 9 * 16, 17 * 9 * 16
PQ_ROWB, PQ_PLANE = 96, 36 * 96


def bf_pixel_quad(b, m):
    """conv_b1fused.cu bf_pixel<true>: (valid, R, C, pc) of lane m of batch b of a unit (CH = 64)."""
    idx = b * 128 + m
    if idx < 272:
        pc = 1 if idx >= 136 else 0
        rem = idx - 136 * pc
        R, C = rem >> 3, rem & 7
    else:
        pc = 1 if idx >= 289 else 0
        R, C = idx - 272 - 17 * pc, 8
    return idx < 306, R, C, pc


def bf_pixel_128(s, pr, b, m):
    """conv_b1fused.cu bf_pixel<false>: (valid, R, C, pc) for strip s, row parity pr (CH = 128)."""
    valid = True
    if b == 0:
        rr, C, pc = m >> 3, (m & 7) + (1 - s), s
    else:
        idx = (b - 1) * 128 + m
        valid = idx < 144
        rr = idx // 9
        C, pc = idx - rr * 9, 1 - s
    return valid, rr + (0 if pr else 1), C, pc


def test_quad_map_tiles_both_planes_once():
    cells = [bf_pixel_quad(b, m)[1:] for b in range(3) for m in range(128) if bf_pixel_quad(b, m)[0]]
    assert len(cells) == 306 and len(set(cells)) == 306
    assert set(cells) == {(R, C, pc) for pc in range(2) for R in range(17) for C in range(9)}
    # the third batch of a unit holds 50 pixels: lane quadrants 2, 3 skip their TMEM loads (W_SKIP)
    assert [bf_pixel_quad(2, m)[0] for m in (0, 49, 50, 64, 127)] == [True, True, False, False, False]


def test_ch128_map_tiles_the_unit_once():
    for s, pr in itertools.product(range(2), range(2)):
        cells = [bf_pixel_128(s, pr, b, m)[1:] for b in range(3) for m in range(128) if bf_pixel_128(s, pr, b, m)[0]]
        assert len(cells) == 272 and len(set(cells)) == 272                  # 16 x 8 + 16 x 9
        rows = {R for R, _, _ in cells}
        assert rows == set(range(0, 16) if pr else range(1, 17))             # the halo rows ty = 0, 33 are never written
        for R, C, pc in cells:
            tx = 2 * C + pc                                                  # image column - 16 s + 1: 1..16 computed, 0 | 17 = halo
            assert (1 if s == 0 else 0) <= tx <= (17 if s == 0 else 16)
        assert all(not bf_pixel_128(s, pr, 2, m)[0] for m in range(16, 128))  # third batch: 16 pixels, lane quadrant 0 only


def _quad_word(set_, k, s, q, lane):
    """the packed word pw[k] of conv_b1fused.cu's CH = 64 T loop, as (g_off, c_off, omask, flags) before packing"""
    r = set_ + 2 * k
    uu, b = divmod(r, 3)
    pr = 1 - uu
    valid, R, C, pc = bf_pixel_quad(b, q * 32 + lane)
    y, x = 2 * R + pr - 1, 16 * s - 1 + 2 * C + pc
    g_off = pc * PQ_PLANE + (y + 1) * PQ_ROWB + C * 8
    c_off = ((pr * 2 + pc) * 8) * BF_K8 + R * BF_ROW + C * 16
    omask = 0
    for quad in range(4):
        Y, X = 32 * (quad >> 1) + y, 32 * (quad & 1) + x
        if not (0 <= Y < 64 and 0 <= X < 64):
            omask |= 1 << quad
    return valid, g_off, c_off, omask, pc, (y, x)


def test_quad_packed_words_fit_and_stay_in_bounds():
    t_bytes, p_bytes = 4 * 8 * BF_K8, 2 * PQ_PLANE
    for set_, k, s, q, lane in itertools.product(range(2), range(3), range(2), range(4), range(32)):
        valid, g_off, c_off, omask, pc, (y, x) = _quad_word(set_, k, s, q, lane)
        if not valid:
            continue
        assert g_off % 8 == 0 and (g_off >> 3) < 0x3ff                       # 10 bits, all ones = "no pixel"
        assert c_off % 16 == 0 and (c_off >> 4) < (1 << 13)
        assert c_off + 7 * BF_K8 + 16 <= t_bytes                             # the eight channel groups of the cell
        # the nine taps: rows y + 1 .. y + 3 of the patch, entries C (+1) of the own plane and C + pc of the other one
        mid = g_off + (8 - PQ_PLANE if pc else PQ_PLANE)
        for ky in range(3):
            for off in (g_off, mid, g_off + 8):
                a = off + ky * PQ_ROWB
                assert 0 <= a and a + 8 <= p_bytes
                plane, rem = divmod(a, PQ_PLANE)
                row, ent = divmod(rem, PQ_ROWB)
                assert ent + 8 <= 80 and row < 36                            # 10 entries per 96-byte row, 36 rows
        assert -1 <= y <= 32 and 16 * s - 1 <= x <= 16 * s + 16
        # outside-the-image bits: only the tile's outer ring can leave the image, and only on the image's border quadrants
        assert omask == sum(1 << quad for quad in range(4)
                            if (y < 0 and quad < 2) or (y > 31 and quad >= 2) or (x < 0 and not quad & 1) or (x > 31 and quad & 1))


def test_quad_gathers_are_conflict_free_for_8_cell_rows():
    """LDS.64: a half-warp (16 lanes) is served in one wavefront when its 8-byte words fall into 16 distinct bank pairs."""
    for set_, k, s, q in itertools.product(range(2), range(3), range(2), range(4)):
        for half in range(2):
            lanes = [_quad_word(set_, k, s, q, half * 16 + l) for l in range(16)]
            idx0 = (set_ + 2 * k) % 3 * 128 + q * 32 + half * 16
            if idx0 + 15 >= 272 or (idx0 < 136 <= idx0 + 15):                # ninth-cell column / plane boundary: replays accepted
                continue
            pairs = {(w[1] % 128) // 8 for w in lanes}
            assert len(pairs) == 16, (set_, k, s, q, half)
