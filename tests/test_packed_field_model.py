"""Arithmetic model of the packed rows of the mode-2 kernel (csrc/poa_gap_blk.cu, DESIGN.md "Range safety"): two scores
per 32-bit word as biased 16-bit fields. Checks, with numpy on uint32 words, the three facts the kernel relies on:
  1. a plain 32-bit add of `c * 65537` adds c to both fields while both results stay inside [0, 65535];
  2. bit 15 / bit 31 of `a + 0x80008000 - b` are the comparisons `a >= b` of the two fields whenever |a - b| < 32768;
  3. with the kernel's bounds (real cells in [-21920, +3960], garbage never below -27000) every difference the row takes
     satisfies 2, and outside that window the trick does break (so the guards are needed, not decorative).
No GPU, no product code: this is the design's arithmetic, not the implementation (the implementation is A/B-tested on
the GPU against the 32-bit rows and the striped kernel)."""
import numpy as np

BIAS = 32768
K = np.uint32(0x80008000)
REAL_LO, REAL_HI, GARBAGE_LO = -21920, 3960, -27000


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


def test_flag_words_land_in_natural_column_order():
    # bit r <- lo cell of pair r, bit 16 + r <- hi cell: (f >> (15 - r)) & (1 << r | 1 << (16 + r)), as in the kernel
    rng = np.random.default_rng(3)
    lo = rng.integers(0, 2, (1000, 16)).astype(np.uint32)
    hi = rng.integers(0, 2, (1000, 16)).astype(np.uint32)
    acc = np.zeros(1000, dtype=np.uint32)
    for r in range(16):
        f = (lo[:, r] << np.uint32(15)) | (hi[:, r] << np.uint32(31)) | np.uint32(0x12340123)  # flags + garbage below them
        f &= ~np.uint32(0x80008000) | (lo[:, r] << np.uint32(15)) | (hi[:, r] << np.uint32(31))
        acc |= (f >> np.uint32(15 - r)) & np.uint32((1 << r) | (1 << (16 + r)))
    want = np.zeros(1000, dtype=np.uint32)
    for r in range(16):
        want |= (lo[:, r] << np.uint32(r)) | (hi[:, r] << np.uint32(16 + r))
    assert np.array_equal(acc, want)
