"""Arithmetic model of the packed rows of the mode-2 kernel (csrc/poa_gap_blk.cu, DESIGN.md "Range safety"): two scores
per 32-bit word as biased 16-bit fields. Checks, with numpy on uint32 words, the three facts the kernel relies on:
  1. a plain 32-bit add of `c * 65537` adds c to both fields while both results stay inside [0, 65535];
  2. bit 15 / bit 31 of `a + 0x80008000 - b` are the comparisons `a >= b` of the two fields whenever |a - b| < 32768;
  3. with the kernel's bounds (real cells in [-21920, +4920], garbage never below -27000) every difference the row takes
     satisfies 2, and outside that window the trick does break (so the guards are needed, not decorative).
No GPU, no product code: this is the design's arithmetic, not the implementation (the implementation is A/B-tested on
the GPU against the 32-bit rows and the striped kernel)."""
import numpy as np

BIAS = 32768
K = np.uint32(0x80008000)
REAL_LO, REAL_HI, GARBAGE_LO = -21920, 4920, -27000


def pack(lo, hi):
    return ((hi + BIAS).astype(np.uint32) << np.uint32(16)) | (lo + BIAS).astype(np.uint32)


def unpack(w):
    return (w & np.uint32(0xFFFF)).astype(np.int64) - BIAS, (w >> np.uint32(16)).astype(np.int64) - BIAS


def add2(c):
    return np.uint32((c * 65537) & 0xFFFFFFFF)


def test_adding_a_constant_to_both_fields_is_one_32_bit_add():
    rng = np.random.default_rng(1)
    lo = rng.integers(GARBAGE_LO, REAL_HI + 1, 200000)
    hi = rng.integers(GARBAGE_LO, REAL_HI + 1, 200000)
    w = pack(lo, hi)
    for c in (-60, -30, -8, -2, -1, 0, 1, 2, 30, 60, -2000):
        rl, rh = unpack(w + add2(c))
        assert np.array_equal(rl, lo + c) and np.array_equal(rh, hi + c)


def test_flags_are_the_sign_positions_of_a_biased_difference():
    rng = np.random.default_rng(2)
    n = 400000
    # operands as the row sees them: real cells and garbage (padding, forced first-column entries) in the same word
    a_lo = rng.integers(GARBAGE_LO, REAL_HI + 1, n)
    a_hi = rng.integers(GARBAGE_LO, REAL_HI + 1, n)
    b_lo = rng.integers(GARBAGE_LO, REAL_HI + 1, n)
    b_hi = rng.integers(GARBAGE_LO, REAL_HI + 1, n)
    # make ties and near-ties frequent: they are what the tie-breaking rules of the reference depend on
    tie = rng.random(n) < 0.3
    b_lo = np.where(tie, a_lo + rng.integers(-1, 2, n), b_lo)
    b_hi = np.where(rng.random(n) < 0.3, a_hi + rng.integers(-1, 2, n), b_hi)
    b_lo = np.clip(b_lo, GARBAGE_LO, REAL_HI)
    b_hi = np.clip(b_hi, GARBAGE_LO, REAL_HI)
    assert (REAL_HI - GARBAGE_LO) < 32768
    f = pack(a_lo, a_hi) + K - pack(b_lo, b_hi)
    assert np.array_equal((f >> np.uint32(15)) & np.uint32(1), (a_lo >= b_lo).astype(np.uint32))
    assert np.array_equal((f >> np.uint32(31)) & np.uint32(1), (a_hi >= b_hi).astype(np.uint32))
    # the strict forms the kernel uses: y > m + o  ==  (y - 1) >= (m + o);  x > m + o  ==  x >= m + o + 1
    o = -4
    fy = pack(a_lo, a_hi) + (K - add2(1)) - pack(b_lo, b_hi)
    assert np.array_equal((fy >> np.uint32(15)) & np.uint32(1), (a_lo > b_lo).astype(np.uint32))
    assert np.array_equal((fy >> np.uint32(31)) & np.uint32(1), (a_hi > b_hi).astype(np.uint32))
    fx = pack(a_lo, a_hi) + (K - add2(o + 1)) - pack(b_lo, b_hi)
    assert np.array_equal((fx >> np.uint32(15)) & np.uint32(1), (a_lo > b_lo + o).astype(np.uint32))
    assert np.array_equal((fx >> np.uint32(31)) & np.uint32(1), (a_hi > b_hi + o).astype(np.uint32))


def test_outside_the_window_a_field_corrupts_its_neighbour():
    # lo difference of -40000 (below -32768): the borrow reaches the hi field and flips a hi tie
    a = pack(np.array([-30000]), np.array([100]))
    b = pack(np.array([10000]), np.array([100]))
    f = a + K - b
    assert int(((f >> np.uint32(31)) & np.uint32(1))[0]) == 0  # wrong: 100 >= 100 is true
    # inside the window the same tie is reported correctly
    a = pack(np.array([-20000]), np.array([100]))
    f = a + K - b
    assert int(((f >> np.uint32(31)) & np.uint32(1))[0]) == 1


def perm(k, C):
    """PlaneFmt<C>::perm of csrc/poa_gap_blk.cu: plane bit of column k of a lane's block of C columns."""
    H, G = C // 2, C // 4
    half = 1 if k >= H else 0
    kk = k - half * H
    return (2 * half + (kk & 1)) * G + (kk >> 1)


def prmt_sign_bytes(f_even, f_odd):
    """prmt.b32 d, f_even, f_odd, 0xfbd9: bytes (f_even.b1, f_odd.b1, f_even.b3, f_odd.b3), each replaced by its sign bit
    replicated over the byte (selector nibbles 9, d, b, f: bit 3 = replicate the msb)."""
    out = np.zeros_like(f_even)
    for j, (w, byte) in enumerate(((f_even, 1), (f_odd, 1), (f_even, 3), (f_odd, 3))):
        sign = (w >> np.uint32(8 * byte + 7)) & np.uint32(1)
        out |= (sign * np.uint32(0xFF)) << np.uint32(8 * j)
    return out


def packed_row_planes(lo, hi, C):
    """Flag accumulation of row16: one byte permute per two cell pairs, AND-OR into bit r/2 of the four byte groups,
    byte groups squeezed to G = C/4 bits for C < 32."""
    H, G = C // 2, C // 4
    n = lo.shape[0]
    acc = np.zeros(n, dtype=np.uint32)
    rng = np.random.default_rng(9)
    f = []
    for r in range(H):
        garbage = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32) & ~np.uint32(0x80008000)
        f.append(garbage | (lo[:, r] << np.uint32(15)) | (hi[:, r] << np.uint32(31)))
    for r in range(0, H, 2):
        acc |= prmt_sign_bytes(f[r], f[r + 1]) & np.uint32(0x01010101 << (r // 2))
    if H < 16:
        gm = np.uint32((1 << G) - 1)
        acc = (acc & gm) | (((acc >> np.uint32(8)) & gm) << np.uint32(G)) | (((acc >> np.uint32(16)) & gm) << np.uint32(2 * G)) | \
              (((acc >> np.uint32(24)) & gm) << np.uint32(3 * G))
    return acc


def test_plane_bit_order_is_a_permutation_and_the_byte_permute_produces_it():
    rng = np.random.default_rng(3)
    for C in (4, 8, 16, 32):
        H = C // 2
        assert sorted(perm(k, C) for k in range(C)) == list(range(C))
        assert perm(0, C) == 0 and perm(C - 1, C) == C - 1   # the first-column fix and the carry to the next lane rely on these
        lo = rng.integers(0, 2, (2000, H)).astype(np.uint32)
        hi = rng.integers(0, 2, (2000, H)).astype(np.uint32)
        acc = packed_row_planes(lo, hi, C)
        want = np.zeros(2000, dtype=np.uint32)
        for r in range(H):
            want |= (lo[:, r] << np.uint32(perm(r, C))) | (hi[:, r] << np.uint32(perm(r + H, C)))
        assert np.array_equal(acc, want), C


def test_path_x_shift_in_plane_order():
    # path_x of column c is the x-flag of column c - 1: the kernel moves every bit to the plane position of the next column
    rng = np.random.default_rng(4)
    for C in (4, 8, 16, 32):
        G = C // 4
        gm = (1 << G) - 1 if G < 32 else 0xFFFFFFFF
        EVEN = gm | (gm << (2 * G))
        ODDIN = ((gm >> 1) << G) | ((gm >> 1) << (3 * G))
        LOLAST = 1 << (2 * G - 1)
        for _ in range(500):
            flags = rng.integers(0, 2, C)
            carry = int(rng.integers(0, 2))
            fx = sum(int(flags[k]) << perm(k, C) for k in range(C))
            nx = ((fx & EVEN) << G) | ((fx & LOLAST) << 1) | carry
            if G > 1:
                nx |= (fx & ODDIN) >> (G - 1)
            nx &= (1 << C) - 1 if C < 32 else 0xFFFFFFFF
            want = carry | sum(int(flags[k - 1]) << perm(k, C) for k in range(1, C))
            assert nx == want, (C, flags, carry)
            assert (fx >> (C - 1)) & 1 == flags[C - 1]   # what the next lane receives as its carry


def test_moving_base_removes_the_per_cell_gap_extension_add():
    # row16 computes a row relative to base + e: y = max(m' + o, y') + e of the previous row's fields is then
    # max(m' + o, y') as it stands (one VIADDMNMX), and the diagonal uses substitution addends s - e
    rng = np.random.default_rng(5)
    n = 100000
    o, e, base = -4, -2, 1234
    m_prev = rng.integers(REAL_LO + 100, REAL_HI - 100, n)
    y_prev = rng.integers(REAL_LO + 100, REAL_HI - 100, n)
    s = rng.choice([2, -4], n)
    field = lambda score, b: score - b + BIAS
    y_new = np.maximum(m_prev + o, y_prev) + e
    d_new = m_prev + s
    nb = base + e
    assert np.array_equal(field(y_new, nb), np.maximum(field(m_prev, base) + o, field(y_prev, base)))
    assert np.array_equal(field(d_new, nb), field(m_prev, base) + (s - e))
    # the Y flag: y' > m' + o  ==  y' + (BIAS2 - (o + 1)) - m' has its sign position set
    w = pack(y_prev - base, y_prev - base) + (K - add2(o + 1)) - pack(m_prev - base, m_prev - base)
    assert np.array_equal((w >> np.uint32(15)) & np.uint32(1), (y_prev > m_prev + o).astype(np.uint32))
