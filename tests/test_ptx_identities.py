"""The arithmetic identities the PTX fast decoder (lzma_b200/csrc/lzgpu_fast2.cuh) rests on, checked against the
reference's formulas (range_decoder.go:57-134, prob update: decompress.go passim) with numpy over random and boundary
inputs.  The PTX itself only runs on the device (GPU tier); what can be wrong in it without any hardware -- the
algebra -- is pinned here.
"""
import numpy as np

M32 = np.uint64(0xFFFFFFFF)
rng = np.random.default_rng(20260)


def u32(x):
    return np.asarray(x, dtype=np.uint64) & M32


def rand_u32(n, lo=0, hi=1 << 32):
    a = rng.integers(lo, hi, n, dtype=np.uint64)
    edge = np.array([lo, hi - 1, (lo + hi) // 2, lo + 1, hi - 2], dtype=np.uint64)
    return np.concatenate([a, edge])


def test_code_update_as_add_min():
    """F2_CSUB: code = min(code, code - bound) (mod 2^32)  ==  code >= bound ? code - bound : code,  for bound < 2^32 and
    any code (in the decoder code < range <= 2^32 - 1 and bound < range)."""
    code = rand_u32(200_000)
    bound = rand_u32(200_000)[rng.permutation(code.size)]
    want = np.where(code >= bound, code - bound, code)
    got = np.minimum(code, u32(code - bound + (np.uint64(1) << np.uint64(32))))
    assert np.array_equal(got, want)


def ref_direct_bits(rng_, code, n, nxt):
    """rangeDecoder.DecodeDirectBits (range_decoder.go:101-134) on python ints; nxt() yields input bytes."""
    res = 0
    for _ in range(n):
        rng_ >>= 1
        code = (code - rng_) & 0xFFFFFFFF
        t = (0 - (code >> 31)) & 0xFFFFFFFF
        code = (code + (rng_ & t)) & 0xFFFFFFFF
        if rng_ < (1 << 24):
            rng_ = (rng_ << 8) & 0xFFFFFFFF
            code = ((code << 8) | nxt()) & 0xFFFFFFFF
        res = ((res << 1) + (t + 1)) & 0xFFFFFFFF
    return rng_, code, res


def f2_direct8(code, r, live=8):
    """F2_DIRECT8 / F2_DIRECT_PART: steps j = 1..8 of code = min(code, code - (live ? r >> j : 0)); a step that changed
    code decided 1."""
    acc = 0
    for j in range(1, 9):
        rj = (r >> j) if j <= live else 0
        t = (code - rj) & 0xFFFFFFFF
        c2 = min(code, t)
        if c2 != code:
            acc |= 1 << (8 - j)
        code = c2
    return code, acc


def f2_direct_bits(rng_, code, n, nxt):
    """The fast decoder's arrangement of the n direct bits (lzgpu_fast2.cuh, distance decode): a first run of g = 8 -
    clz(range) real steps made a full one by pre-shifting the range, runs of 8 after every normalisation, a last
    partial run."""
    res = 0
    g = 8 - (32 - rng_.bit_length())
    assert 1 <= g <= 8
    if n >= g:
        code, acc = f2_direct8(code, (rng_ << (8 - g)) & 0xFFFFFFFF)
        res = acc
        rng_ >>= g
        n -= g
        rng_, code = (rng_ << 8) & 0xFFFFFFFF, ((code << 8) | nxt()) & 0xFFFFFFFF
        while n >= 8:
            code, acc = f2_direct8(code, rng_)
            res = ((res << 8) | acc) & 0xFFFFFFFF
            rng_ >>= 8
            n -= 8
            rng_, code = (rng_ << 8) & 0xFFFFFFFF, ((code << 8) | nxt()) & 0xFFFFFFFF
    if n:
        code, acc = f2_direct8(code, rng_, live=n)
        res = ((res << n) | (acc >> (8 - n))) & 0xFFFFFFFF
        rng_ >>= n
    return rng_, code, res


def test_direct_bits_as_add_min_runs():
    import random
    r = random.Random(5)
    for trial in range(20_000):
        rng0 = r.randrange(1 << 24, 1 << 32) if trial % 7 else r.choice([1 << 24, (1 << 32) - 1, (1 << 31), (1 << 31) - 1, (1 << 25) + 1])
        code0 = r.randrange(0, rng0)
        n = r.randrange(1, 27)
        data = [r.randrange(256) for _ in range(8)]
        it1, it2 = iter(data), iter(data)
        want = ref_direct_bits(rng0, code0, n, lambda: next(it1))
        got = f2_direct_bits(rng0, code0, n, lambda: next(it2))
        assert got == want, (hex(rng0), hex(code0), n)
        assert len(list(it1)) == len(list(it2))      # the same number of input bytes consumed


def test_next_bound_from_the_unnormalised_range():
    """F2_TAHEAD / F2_TSEL: (range >> 11) of the NEXT step taken from the un-normalised range: >> 3 when the
    normalisation is going to shift it left by 8 (range < 2^24), else >> 11."""
    r = rand_u32(200_000, 1 << 13, 1 << 32)
    norm = np.where(r < (1 << 24), u32(r << np.uint64(8)), r)
    want = norm >> np.uint64(11)
    got = np.where(r < (1 << 24), r >> np.uint64(3), r >> np.uint64(11))
    assert np.array_equal(got, want)


def test_probability_update_in_one_formula():
    """F2_UPD: pn = p + (((bit ? 31 : 2048) - p) >> 5) with an arithmetic shift  ==  the reference's update
    (bit 0: p += (2048 - p) >> 5; bit 1: p -= p >> 5), for every 11-bit p."""
    p = np.arange(0, 2048, dtype=np.int64)
    for bit in (0, 1):
        want = p + ((2048 - p) >> 5) if bit == 0 else p - (p >> 5)
        got = p + (((31 if bit else 2048) - p) >> 5)
        assert np.array_equal(got, want)


def test_matched_literal_cell_permutation():
    """The default decoder stores the matched-literal node for (prefix m, match bit b) at cell 0x100 + 2m + b and the plain
    node m at cell m: a bijection with the reference's ((1 + b) << 8) + m and m over the 0x300 cells of a context
    (cells 0, 0x100, 0x101 unused in both), and level i of a matched walk is x = (0x100 | mb) >> (7 - i), u = 2x."""
    seen = set()
    for m in range(1, 256):
        seen.add(m)
        for b in (0, 1):
            seen.add(0x100 + 2 * m + b)
    assert len(seen) == 255 * 3 and max(seen) < 0x300
    for mb in range(256):
        prefix = 1
        for i in range(8):
            b = (mb >> (7 - i)) & 1
            x = (0x100 | mb) >> (7 - i)
            assert x == 2 * prefix + b
            assert 0x100 + x == 0x100 + 2 * prefix + b      # the cell at [u + 512] with u = S + 2x
            prefix = (prefix << 1) | b
