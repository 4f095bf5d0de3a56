"""The PTX of the fast decoder (lzma_b200/csrc/lzgpu_fast2.cuh: the bit ladders that ship) executed on the CPU by a small
PTX interpreter (tests/ptx/interp.py) and compared with the plaintext: isMatch / isRep / length / posSlot / direct bits /
align / literal ladders, matched literals with their mismatch exits, the add-min forms, the normalisation through the
byte-ahead register.  The C++ around the asm blocks (decode_fast2's symbol loop) is restated here in Python, line by
line; the window copy is a plain byte copy.  On the device the same text runs under the GPU tier; here it runs where no
GPU is, which is the only place the rest of the CPU tier cannot reach (the lane emulation runs the careful decoder).
"""
import lzma
import struct

import pytest

from lzma_b200 import corpus as K
from ptx import interp as I

M = 0xFFFFFFFF


class Lay:
    def __init__(self, pb_bits):
        self.PB = pb_bits
        nps = 1 << pb_bits
        self.IS_MATCH = 0
        self.IS_REP0_LONG = 12 * nps
        self.REP4 = 24 * nps
        self.LEN_LOW, self.LEN_MID, self.LEN_HIGH = 8, 8 + 8 * nps, 8 + 16 * nps
        self.LEN_SIZE = 8 + 16 * nps + 256
        self.LEN0 = self.REP4 + 48
        self.LEN1 = self.LEN0 + self.LEN_SIZE
        self.POS_SLOT = self.LEN1 + self.LEN_SIZE
        self.POS_DEC = self.POS_SLOT + 256
        self.ALIGN = self.POS_DEC + 128
        self.FIXED = self.ALIGN + 16
        self.LIT = self.FIXED


INV = {
    "BIT": "F2_BIT(D, PV, A, BIT)",
    "LIT_PRE": "F2_LIT_PRE(D, OUT, S, P0, PLO, PHI)",
    "LIT": "F2_LIT(D, OUT, S, MB, MATCHED)",
    "ISREP_LEN": "F2_ISREP_LEN(D, OUT, SLEN, SLOW, AREP, PREP)",
    "LEN": "F2_LEN(D, OUT, SLEN, SLOW)",
    "TREE6": "F2_TREE6(D, OUT, BASE)",
    "TREE4": "F2_TREE4(D, OUT, BASE, P0, PLO, PHI)",
    "SHIFT8": "F2_SHIFT8(D)",
    "DIRECT8": "F2_DIRECT8(CODE, ACC, R)",
    "DIRECT_PART": "F2_DIRECT_PART(CODE, ACC, R, KK)",
}


@pytest.fixture(scope="module")
def blocks():
    return I.extract(INV)


class Fast:
    """decode_fast2 (lzgpu_fast2.cuh) on one lane.  `out` is the window; returns when the input or the wanted number of
    output bytes is nearly used up, or at the end-of-stream marker."""

    def __init__(self, blocks, stream, want_out):
        self.b = blocks
        props = stream[0]
        self.lc, rest = props % 9, props // 9
        self.pb, self.lp = rest // 5, rest % 5
        self.dict_size = max(struct.unpack("<I", stream[1:5])[0], 4096)
        payload = stream[13:]
        assert payload[0] == 0
        self.Y = Lay(2 if self.pb <= 2 else 4)
        ncells = self.Y.FIXED + (0x300 << (self.lc + self.lp))
        self.sP = 0
        self.sL = 2 * self.Y.LIT
        self.sIn = (2 * ncells + 15) & ~15
        # bounds: stores and 16-bit loads inside the tables only; input bytes from the staged input (+ the byte ahead) only
        self.sm = I.Shared(self.sIn + len(payload) + 64, table_end=2 * ncells, in_lo=self.sIn, in_hi=self.sIn + len(payload) - 5 + 1)
        for c in range(ncells):
            self.sm.st(2 * c, 2, 1024)
        self.sm.b[self.sIn:self.sIn + len(payload) - 5] = payload[5:]
        self.range, self.code = 0xFFFFFFFF, int.from_bytes(payload[1:5], "big")
        self.ips = self.sIn
        self.nb = self.sm.b[self.ips]
        self.lims = self.sIn + len(payload) - 5 - 41 - 1
        self.want_out = want_out
        self.out = bytearray()
        self.rep = [0, 0, 0, 0]
        self.state = 0
        self.wpos, self.full = 0, 0
        self.lp_mask, self.pos_mask = (1 << self.lp) - 1, (1 << self.pb) - 1
        self.eos = False
        self.counters = {}

    # ---- asm blocks: operand lists in the order of the macro's output then input constraints
    def io(self):
        return [self.range, self.code, self.nb, self.ips]

    def set_io(self, v):
        self.range, self.code, self.nb, self.ips = v[0], v[1], v[2], v[3]

    def call(self, name, extra_out, ins):
        v = I.run(self.b[name], self.io() + [0] * extra_out + list(ins), self.sm, self.counters)
        self.set_io(v)
        return v[4] if extra_out else None

    def bit(self, pv, a):
        return self.call("BIT", 1, [a, pv])

    def lds16(self, a):
        return self.sm.ld(a, 2)

    class Rejected(Exception):
        """the C++ around the blocks would leave the fast decoder here (bad distance, rep on an empty window)"""

    def next_ctx(self):
        Y = self.Y
        self.pos_state = self.wpos & self.pos_mask
        self.a_im = self.sP + 2 * Y.IS_MATCH + 2 * ((self.state << Y.PB) + self.pos_state)
        self.a_rep = self.sP + 2 * Y.REP4 + 8 * self.state
        self.p_im, self.p_rep = self.lds16(self.a_im), self.lds16(self.a_rep)

    def bump(self, n):
        self.wpos += n
        if self.wpos >= self.dict_size:
            self.wpos -= self.dict_size
            self.full = 1

    def direct(self, name, r, k=None):
        v = I.run(self.b[name], [self.code, 0, r & M] + ([k] if k is not None else []), self.sm, self.counters)
        self.code = v[0]
        return v[1]

    def run(self):
        Y, sP = self.Y, self.sP
        self.next_ctx()
        pl_valid = False
        while True:
            if self.ips > self.lims or len(self.out) + 274 > self.want_out:
                return
            bit = self.bit(self.p_im, self.a_im)
            if bit == 0:                                           # literal
                if pl_valid:
                    self.state = max(self.state, 3) - 3
                    self.bump(1)
                    self.next_ctx()
                    sym = self.call("LIT_PRE", 1, [pl_S, pl_p, pl_lo, pl_hi])
                else:
                    prevb = self.out[-1] if self.out else 0
                    hist = len(self.out)
                    matchb = self.out[-(self.rep[0] + 1)] if self.rep[0] + 1 <= hist else 0
                    S = self.sL + 0x600 * (((self.wpos & self.lp_mask) << self.lc) + (prevb >> (8 - self.lc)))
                    matched = 1 if self.state >= 7 else 0
                    ns = max(self.state, 3) - 3
                    if self.state >= 10:
                        ns -= 3
                    self.state = ns
                    self.bump(1)
                    self.next_ctx()
                    sym = self.call("LIT", 1, [S, 0x100 | matchb, matched])
                self.out.append(sym & 0xFF)
                assert sym < 256
                pl_S = self.sL + 0x600 * (((self.wpos & self.lp_mask) << self.lc) + (sym >> (8 - self.lc)))
                pl_p, pl_lo, pl_hi = self.lds16(pl_S + 2), self.lds16(pl_S + 4), self.lds16(pl_S + 6)
                pl_valid = True
                continue
            pl_valid = False
            state2 = (self.a_im - sP - 2 * Y.IS_MATCH) >> 1
            a_rep_cur, pos_state_cur = self.a_rep, self.pos_state
            n_mid, n_hi = 2 * (Y.LEN_MID - Y.LEN_LOW), 2 * Y.LEN_HIGH
            ln = self.call("ISREP_LEN", 1, [sP + 2 * Y.LEN0, sP + 2 * (Y.LEN0 + Y.LEN_LOW) + 16 * pos_state_cur, self.a_rep, self.p_rep, n_mid, n_hi])
            if ln != 0xFFFFFFFF:                                   # simple match
                self.rep[3], self.rep[2], self.rep[1] = self.rep[2], self.rep[1], self.rep[0]
                self.state = 7 if self.state < 7 else 10
                len_state = min(ln, 3)
                ln += 2
                wpos0, full0 = self.wpos, self.full
                self.bump(ln)
                self.next_ctx()
                slot = self.call("TREE6", 1, [sP + 2 * Y.POS_SLOT + (len_state << 7)]) - 64
                if slot < 4:
                    self.rep[0] = slot
                else:
                    nd = (slot >> 1) - 1
                    dist = (2 | (slot & 1)) << nd
                    if slot < 14:
                        tb = sP + 2 * (Y.POS_DEC + dist - 4)
                        m, v = 1, 0
                        for i in range(nd):
                            a = tb + 2 * m
                            bit = self.bit(self.lds16(a), a)
                            m = (m << 1) | bit
                            v |= bit << i
                        dist += v
                    else:
                        al0, al2, al3 = (self.lds16(sP + 2 * Y.ALIGN + o) for o in (2, 4, 6))
                        res, n = 0, nd - 4
                        g = 8 - (32 - self.range.bit_length())
                        if n >= g:
                            res = self.direct("DIRECT8", (self.range << (8 - g)) & M)
                            self.range >>= g
                            n -= g
                            self.call("SHIFT8", 0, [])
                            while n >= 8:
                                res = ((res << 8) | self.direct("DIRECT8", self.range)) & M
                                self.range >>= 8
                                n -= 8
                                self.call("SHIFT8", 0, [])
                        if n:
                            acc = self.direct("DIRECT_PART", self.range, n)
                            res = ((res << n) | (acc >> (8 - n))) & M
                            self.range >>= n
                        dist = (dist + (res << 4)) & M
                        m = self.call("TREE4", 1, [sP + 2 * Y.ALIGN, al0, al2, al3])
                        dist = (dist + (int("{:032b}".format(m)[::-1], 2) >> 28)) & M
                    self.rep[0] = dist
                if self.rep[0] >= (self.dict_size if full0 else wpos0):
                    if self.rep[0] == 0xFFFFFFFF:
                        self.eos = True
                        return
                    raise Fast.Rejected("distance beyond the window: %#x" % self.rep[0])
            else:                                                  # rep match
                if not (self.wpos or self.full):
                    raise Fast.Rejected("rep match on an empty window")
                short_rep = False
                a = a_rep_cur + 2
                bit = self.bit(self.lds16(a), a)
                if bit == 0:
                    a = sP + 2 * Y.IS_REP0_LONG + 2 * state2
                    short_rep = self.bit(self.lds16(a), a) == 0
                else:
                    a = a_rep_cur + 4
                    if self.bit(self.lds16(a), a) == 0:
                        self.rep[0], self.rep[1] = self.rep[1], self.rep[0]
                    else:
                        a = a_rep_cur + 6
                        if self.bit(self.lds16(a), a) == 0:
                            self.rep[0], self.rep[1], self.rep[2] = self.rep[2], self.rep[0], self.rep[1]
                        else:
                            self.rep[0], self.rep[1], self.rep[2], self.rep[3] = self.rep[3], self.rep[0], self.rep[1], self.rep[2]
                if short_rep:
                    self.state = 9 if self.state < 7 else 11
                    ln = 1
                else:
                    ln = self.call("LEN", 1, [sP + 2 * Y.LEN1, sP + 2 * (Y.LEN1 + Y.LEN_LOW) + 16 * pos_state_cur, n_mid, n_hi]) + 2
                    self.state = 8 if self.state < 7 else 11
                self.bump(ln)
                self.next_ctx()
            d = self.rep[0] + 1
            if d > len(self.out):
                raise Fast.Rejected("rep distance beyond the window")
            for _ in range(ln):
                self.out.append(self.out[-d])


def _check(blocks, plain, want_out, **kw):
    stream = K.compress_alone(plain, **kw)
    f = Fast(blocks, stream, want_out)
    f.run()
    n = len(f.out)
    assert n >= 0.9 * min(want_out, len(plain)) or f.eos, n       # (it stops 41 input bytes / 274 output bytes short of the end)
    assert bytes(f.out) == plain[:n]
    return f


def test_text_through_the_ptx_ladders(blocks):
    plain = K.text_block(77, 48 << 10)
    f = _check(blocks, plain, len(plain))
    assert f.counters["steps"] > 100_000


def test_mixed_and_incompressible_data(blocks):
    _check(blocks, K.mixed_block(78, 40 << 10), 40 << 10)
    _check(blocks, K.random_block(79, 6 << 10), 6 << 10)


@pytest.mark.parametrize("lc,lp,pb", [(0, 0, 0), (3, 0, 2), (1, 2, 1), (4, 0, 4), (0, 4, 3), (2, 1, 0)])
def test_properties(blocks, lc, lp, pb):
    _check(blocks, K.mixed_block(80 + lc + 3 * lp + 7 * pb, 20 << 10), 20 << 10, lc=lc, lp=lp, pb=pb)


def test_long_distances_and_repeats(blocks):
    """far matches (many direct bits: full 8-bit runs after the first), rep matches, short reps"""
    import random
    r = random.Random(9)
    base = bytes(r.randrange(256) for _ in range(3000))
    plain = bytearray()
    while len(plain) < 160_000:
        plain += base[: r.randrange(20, 400)]
        plain += bytes(r.randrange(256) for _ in range(r.randrange(0, 6)))
        if r.random() < 0.05:
            plain += bytes(r.randrange(256) for _ in range(2500))
    _check(blocks, bytes(plain[:160_000]), 160_000, dict_size=1 << 20)


def test_hostile_streams_stay_inside_the_tables(blocks):
    """compute-sanitizer is closed on this pool; the interpreter's shared memory has bounds instead.  Damaged streams
    -- flipped, overwritten, inserted bytes inside the range-coded part -- drive the ladders with garbage: every store
    and every 16-bit load must stay inside the probability tables, every input byte inside the staged input, whatever
    is decoded (the C++ around the blocks rejects bad distances; wrong bytes are what a damaged stream gives)."""
    import random
    r = random.Random(4)
    plains = [K.text_block(91, 24 << 10), K.mixed_block(92, 24 << 10), K.random_block(93, 4 << 10)]
    n_run = 0
    for plain in plains:
        for lc, lp, pb in ((3, 0, 2), (0, 4, 4), (4, 0, 0)):
            good = K.compress_alone(plain, lc=lc, lp=lp, pb=pb)
            for _ in range(3):
                s = bytearray(good)
                for _ in range(r.randrange(1, 4)):
                    kind, pos = r.randrange(3), r.randrange(18, len(s) - 8)
                    if kind == 0:
                        s[pos] ^= 1 << r.randrange(8)
                    elif kind == 1:
                        s[pos:pos + r.randrange(1, 6)] = bytes(r.randrange(256) for _ in range(r.randrange(1, 6)))
                    else:
                        s[pos:pos] = bytes(r.randrange(256) for _ in range(r.randrange(1, 4)))
                f = Fast(blocks, bytes(s), len(plain))
                try:
                    f.run()
                except Fast.Rejected:
                    pass
                n_run += 1
    assert n_run == 27
