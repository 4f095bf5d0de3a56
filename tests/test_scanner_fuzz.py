"""Host LZMA2 scanner (lzgpu_scan_lzma2 = Reader2.startChunk's framing rules, reader2.go:100-214, SURVEY
Appendix C) against an independent Python model on random chunk sequences, and on random bytes: the units must
tile the input and the output exactly, start only where nothing is inherited, and carry the properties in force."""
import random

from lzma_b200 import _lib as L
from lzma_b200 import batch as B


def _chunk(rng, ctrl_kind, props=None):
    """One chunk with a dummy payload: returns (bytes, control, uncompressed size, props byte or None)."""
    if ctrl_kind in (1, 2):
        usz = rng.randrange(1, 70_000) if rng.random() < 0.2 else rng.randrange(1, 3000)
        usz = min(usz, 65_536)
        return bytes([ctrl_kind, (usz - 1) >> 8, (usz - 1) & 0xFF]) + rng.randbytes(usz), ctrl_kind, usz, None
    usz = rng.randrange(1, 1 << 21) if rng.random() < 0.3 else rng.randrange(1, 5000)
    csz = rng.randrange(1, 65_537) if rng.random() < 0.1 else rng.randrange(1, 2000)
    c = ctrl_kind | ((usz - 1) >> 16)
    h = bytes([c, ((usz - 1) >> 8) & 0xFF, (usz - 1) & 0xFF, (csz - 1) >> 8, (csz - 1) & 0xFF])
    if ctrl_kind >= 0xC0:
        h += bytes([props])
    return h + rng.randbytes(csz), c, usz, props if ctrl_kind >= 0xC0 else None


def _model(chunks, dict_size):
    """Expected units [(in_off, in_len, out_off, out_cap, lc, lp, pb, lit_bits, fresh)] per Appendix C."""
    units = []
    pos = out = 0
    props = 0
    seen_lzma = False
    cur = None

    def open_unit():
        p = props if props < 225 else 0
        return {"in_off": pos, "out_off": out, "lc": p % 9, "lp": (p // 9) % 5, "pb": p // 45, "lit_bits": 0,
                "fresh": not seen_lzma}

    cur = open_unit()
    for k, (raw, c, usz, pr) in enumerate(chunks):
        kind = c if c < 0x80 else c & 0xE0
        starts = False
        if (kind == 1 or kind == 0xE0) and pos != cur["in_off"]:
            starts = True
            if kind == 1:                       # nothing may be inherited by the first LZMA chunk that follows
                for raw2, c2, _, _ in chunks[k + 1:]:
                    k2 = c2 if c2 < 0x80 else c2 & 0xE0
                    if k2 >= 0x80:
                        starts = k2 >= 0xA0
                        break
                    if k2 == 1:
                        break
        if starts:
            cur["in_len"] = pos - cur["in_off"]
            cur["out_cap"] = out - cur["out_off"]
            units.append(cur)
            cur = open_unit()
        if pr is not None:
            props = pr
        if kind >= 0x80:
            if props < 225:
                cur["lit_bits"] = max(cur["lit_bits"], props % 9 + (props // 9) % 5)
            seen_lzma = True
        pos += len(raw)
        out += usz
    cur["in_len"] = pos + 1 - cur["in_off"]     # the terminator belongs to the last unit
    cur["out_cap"] = out - cur["out_off"]
    units.append(cur)
    return units, out


def test_scanner_matches_model_on_random_chunk_sequences():
    rng = random.Random(11)
    for it in range(300):
        chunks = []
        n = rng.randrange(1, 40)
        for k in range(n):
            if k == 0:
                kind = rng.choice([1, 0xE0])
            else:
                kind = rng.choice([1, 2, 2, 0x80, 0x80, 0xA0, 0xC0, 0xE0, 0xE0])
            chunks.append(_chunk(rng, kind, rng.randrange(225)))
        data = b"".join(c[0] for c in chunks) + b"\x00"
        want, total = _model(chunks, 1 << 20)
        units, got_total, sst = B.scan_lzma2(data, 1 << 20)
        assert sst == L.OK and got_total == total, it
        assert len(units) == len(want), (it, len(units), len(want))
        for u, w in zip(units, want):
            assert (u.in_off, u.in_len, u.out_off, u.out_cap) == (w["in_off"], w["in_len"], w["out_off"], w["out_cap"]), it
            assert (u.lc, u.lp, u.pb, u.lit_bits) == (w["lc"], w["lp"], w["pb"], w["lit_bits"]), it
            assert bool(u.flags & L.UF_LZMA2_FRESH) == w["fresh"], it
            assert u.kind == L.KIND_LZMA2_GROUP and u.dict_size == 1 << 20
        assert units[-1].flags & L.UF_LZMA2_LAST and not any(u.flags & L.UF_LZMA2_LAST for u in units[:-1])
        # truncated anywhere: still tiles what is there, reports UNEXPECTED_EOF unless cut right after a full chunk + terminator
        cut = rng.randrange(0, len(data))
        units, _, sst = B.scan_lzma2(data[:cut], 1 << 20)
        assert sst == L.UNEXPECTED_EOF
        assert units[0].in_off == 0 and sum(u.in_len for u in units) == cut
        for a, b in zip(units, units[1:]):
            assert a.in_off + a.in_len == b.in_off and a.out_off + a.out_cap == b.out_off


def test_scanner_on_random_bytes_stays_in_bounds():
    rng = random.Random(12)
    for it in range(500):
        data = rng.randbytes(rng.randrange(0, 400))
        units, total, sst = B.scan_lzma2(data, rng.choice([0, 4096, 1 << 20]))
        assert len(units) >= 1 and sst in (L.OK, L.UNEXPECTED_EOF)
        assert units[0].in_off == 0 and sum(u.in_len for u in units) == len(data)
        for a, b in zip(units, units[1:]):
            assert a.in_off + a.in_len == b.in_off and a.out_off + a.out_cap == b.out_off
        assert units[-1].out_off + units[-1].out_cap == total
