"""CPU: the integer shortcuts K1 (csrc/preprocess.cu) takes are identities of the oracle's
arithmetic.  Each test restates one shortcut in numpy and checks it against
oracle/preprocess.py over the whole input domain, so a change of either side shows up here
without a GPU (the GPU parity tests then check the kernel itself)."""
import numpy as np

from oracle import preprocess as P


def test_luma_quotient_from_one_multiply_high():
    # preprocess.cu luma_8px: h = (8 s * ceil(2^48 / 80000)) >> 32 for s = 2126 R + 7152 G + 722 B;
    # byte 2 of h is s // 10000 and the low 16 bits are zero exactly when 10000 divides s
    s = np.arange(0, 2126 * 255 + 7152 * 255 + 722 * 255 + 1, dtype=np.uint64)
    m = np.uint64(3518437209)
    assert int(m) == -(-(1 << 48) // 80000)
    h = ((s * np.uint64(8)) * m) >> np.uint64(32)
    assert int(h.max()) < 1 << 24                                   # quotient fits byte 2
    assert np.array_equal(h >> np.uint64(16), s // np.uint64(10000))
    assert np.array_equal((h & np.uint64(0xFFFF)) == 0, s % np.uint64(10000) == 0)
    # the 8x-scaled coefficients are 16-bit (dp2a operands)
    assert max(8 * 2126, 8 * 7152, 8 * 722) < 1 << 16


def test_luma_correction_only_on_exact_multiples_and_only_minus_one():
    # the kernel subtracts the bitmap bit only where 10000 | s; everywhere else truncation of the
    # reference's float64 expression equals the integer floor
    r, g = np.meshgrid(np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8), indexing="ij")
    for b in (0, 1, 24, 57, 128, 152, 254, 255):
        t = np.stack([r, g, np.full_like(r, b)], axis=-1)
        ti = t.astype(np.int64)
        s = 2126 * ti[..., 0] + 7152 * ti[..., 1] + 722 * ti[..., 2]
        ref, flo = P.luma_truncate(t).astype(int), P.luma_int_floor(t).astype(int)
        assert np.array_equal(ref[s % 10000 != 0], flo[s % 10000 != 0])
        d = flo - ref
        assert d.min() >= 0 and d.max() <= 1


def test_horizontal_taps_are_periodic_with_compile_time_offsets():
    # preprocess.cu phase B: sx(d) = (80 d + 19) // 42, coefficient pairs repeat every 21 outputs
    # <-> 40 source pixels, and a period's taps lie inside its own 40 source bytes
    sx, c0, c1 = P.cv2_linear_taps(160, 84)
    d = np.arange(84)
    assert np.array_equal(sx, (80 * d + 19) // 42)
    assert np.array_equal(sx[21:], sx[:-21] + 40)
    assert np.array_equal(c0[21:], c0[:-21]) and np.array_equal(c1[21:], c1[:-21])
    assert sx[:21].min() == 0 and (sx[:21] + 1).max() == 39
    assert int(c0.max()) <= 2048 and int(c1.max()) <= 2048 and int(c0.min()) >= 0 and int(c1.min()) >= 0


def test_vertical_taps_give_each_warp_two_exclusive_runs():
    # warp w owns output rows 4w+1 .. 4w+4 (warp 20: 81, 82, 83, 0): their source rows are
    # 10w+3 .. 10w+6 and 10w+8 .. 10w+11 (warp 20: 203-206, 208-209 and 0-1), no row shared
    sy, b0, b1 = P.cv2_linear_taps(210, 84)
    seen = set()
    for w in range(21):
        rows = []
        for ry in range(4):
            dy = (4 * w + 1 + ry) % 84
            rows += [int(sy[dy]), int(sy[dy]) + 1]
        want = (list(range(10 * w + 3, 10 * w + 7)) + list(range(10 * w + 8, 10 * w + 12))
                if w < 20 else [203, 204, 205, 206, 208, 209, 0, 1])
        assert rows == want, w
        assert not (seen & set(rows))
        seen |= set(rows)
    assert len(seen) == 168


def test_packed_vertical_pass_equals_cv2_formula():
    # phase B packs two (b * (h >> 4)) >> 16 terms per 32-bit word and adds the partner row's word
    # plus 0x00020002 before the >> 2: no carry may cross the 16-bit halves
    rng = np.random.default_rng(3)
    y = rng.integers(0, 256, (2, 160)).astype(np.int64)
    y[:, :8] = 255                                                  # the largest sums
    sx, c0, c1 = P.cv2_linear_taps(160, 84)
    for (b0, b1) in ((512, 1536), (1536, 512)):
        h0 = y[0, sx] * c0 + y[0, sx + 1] * c1
        h1 = y[1, sx] * c0 + y[1, sx + 1] * c1
        want = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2
        t0, t1 = (b0 * (h0 >> 4)) >> 16, (b1 * (h1 >> 4)) >> 16
        assert int((b0 * (h0 >> 4)).max()) < 1 << 26
        w0 = t0[0::2] | (t0[1::2] << 16)
        w1 = t1[0::2] | (t1[1::2] << 16)
        s = (w0 + w1 + 0x00020002) & 0xFFFFFFFF
        got = np.empty(84, np.int64)
        got[0::2] = (s >> 2) & 0xFF
        got[1::2] = (s >> 18) & 0xFF
        assert np.array_equal(got, want)
        assert int(want.max()) <= 255


def test_store_address_map_covers_the_blocked_plane_once():
    # phase B: lane (ry, s, m) of warp w stores outputs x = 21 m + s + 2k of row dy = (4w+1+ry) % 84
    # at addr[k & 1] + 16 (k / 2), addr[] built from q0 = m + s; that must be the 4x4-blocked
    # position ((dy/4)*21 + x/4)*16 + (dy%4)*4 + x%4 of src/history.py, each byte exactly once
    seen = np.zeros(7056, np.int32)
    for w in range(21):
        for lane in range(32):
            ry, s, m = lane >> 3, (lane >> 2) & 1, lane & 3
            dy = (4 * w + 1 + ry) % 84
            q0 = m + s
            base = ((dy >> 2) * 21 + 5 * m) * 16 + (dy & 3) * 4
            addr = [base + (q0 >> 2) * 16 + (q0 & 3), base + ((q0 + 2) >> 2) * 16 + ((q0 + 2) & 3)]
            for k in range(11):
                if not (k < 10 or s == 0):
                    continue
                x = 21 * m + s + 2 * k
                a = addr[k & 1] + 16 * (k >> 1)
                assert a == ((dy // 4) * 21 + x // 4) * 16 + (dy % 4) * 4 + x % 4, (w, lane, k)
                seen[a] += 1
    assert seen.min() == 1 and seen.max() == 1


def test_phase_b_byte_offsets_stay_inside_the_lanes_ten_words():
    # tap r of a period reads bytes sx, sx+1 of the lane's 40-byte segment: word k = sx / 4 and,
    # for byte offset 3, word k + 1 -- never past word 9
    for r in range(21):
        sx = (80 * r + 19) // 42
        k, o = sx >> 2, sx & 3
        assert sx + 1 <= 39
        assert k + (1 if o == 3 else 0) <= 9
