"""Seeded synthetic corpora for the parity tests and bench.py.

The reference is decoder-only, so compressed inputs come from liblzma (Python's
``lzma`` module).  Everything here is host-side data preparation; none of it is
on the decode path.
"""
from __future__ import annotations

import lzma
import os
import struct
from concurrent.futures import ProcessPoolExecutor

import numpy as np

UNKNOWN_SIZE = (1 << 64) - 1
_VOCAB_WORDS = 8192


def _vocab(seed: int = 12345):
    """8192 pseudo-words (2..12 lower-case letters, English-like letter weights)."""
    rng = np.random.default_rng(seed)
    letters = np.frombuffer(b"etaoinshrdlcumwfgypbvkjxqz", dtype=np.uint8)
    w = 1.0 / np.arange(1, 27) ** 0.8
    w /= w.sum()
    lens = np.clip(rng.poisson(5.0, _VOCAB_WORDS) + 1, 2, 12).astype(np.int64)
    maxlen = 14
    table = np.zeros((_VOCAB_WORDS, maxlen), dtype=np.uint8)
    for i in range(_VOCAB_WORDS):
        table[i, :lens[i]] = rng.choice(letters, size=lens[i], p=w)
    return table, lens


_VOC = None


def text_block(seed: int, size: int) -> bytes:
    """Text-like block: Zipf(1/rank) words over an 8192-word vocabulary, with
    sentence punctuation, capitals and newlines.  Stream i uses seed i."""
    global _VOC
    if _VOC is None:
        _VOC = _vocab()
    table, lens = _VOC
    rng = np.random.default_rng(1_000_003 * (seed + 1))
    p = 1.0 / np.arange(1, _VOCAB_WORDS + 1)
    p /= p.sum()
    n_words = size // 4 + 64
    ids = rng.choice(_VOCAB_WORDS, size=n_words, p=p)
    wl = lens[ids]
    # separator after each word: ' ' mostly, '. ' / ', ' / '\n' sometimes
    r = rng.random(n_words)
    sep_kind = np.where(r < 0.07, 1, np.where(r < 0.14, 2, np.where(r < 0.16, 3, 0)))  # 1 '. ' 2 ', ' 3 '.\n'
    sep_len = np.where(sep_kind == 0, 1, 2)
    tot = wl + sep_len
    ends = np.cumsum(tot)
    n_keep = int(np.searchsorted(ends, size, side="left")) + 1
    ids, wl, sep_kind, sep_len, tot, ends = (a[:n_keep] for a in (ids, wl, sep_kind, sep_len, tot, ends))
    starts = ends - tot
    out = np.full(int(ends[-1]), 0x20, dtype=np.uint8)
    # scatter the word letters
    row = np.repeat(np.arange(n_keep), wl)
    col = np.arange(int(wl.sum())) - np.repeat(np.cumsum(wl) - wl, wl)
    out[np.repeat(starts, wl) + col] = table[ids[row], col]
    # separators
    sp = starts + wl
    out[sp[sep_kind == 1]] = ord(".")
    out[sp[sep_kind == 2]] = ord(",")
    k3 = sep_kind == 3
    out[sp[k3]] = ord(".")
    out[sp[k3] + 1] = ord("\n")
    # capital after a sentence end
    cap = np.zeros(n_keep, dtype=bool)
    cap[1:] = (sep_kind[:-1] == 1) | (sep_kind[:-1] == 3)
    cap[0] = True
    out[starts[cap]] -= 32
    return out[:size].tobytes()


def random_block(seed: int, size: int) -> bytes:
    return np.random.default_rng(7_000_001 * (seed + 1)).integers(0, 256, size, dtype=np.uint8).tobytes()


def mixed_block(seed: int, size: int) -> bytes:
    """Runs, short-period repeats, text and noise: exercises overlapping copies,
    long matches (273) and rep0..3."""
    rng = np.random.default_rng(9_000_011 * (seed + 1))
    parts, n = [], 0
    while n < size:
        k = int(rng.integers(0, 5))
        ln = int(rng.integers(1, 2000))
        if k == 0:
            b = bytes([int(rng.integers(0, 256))]) * ln
        elif k == 1:
            per = int(rng.integers(2, 9))
            b = (bytes(rng.integers(0, 256, per, dtype=np.uint8)) * (ln // per + 1))[:ln]
        elif k == 2:
            b = text_block(int(rng.integers(0, 1 << 30)), ln)
        elif k == 3:
            b = bytes(rng.integers(0, 256, ln, dtype=np.uint8))
        else:  # structured records: counters + padding
            rec = np.zeros((ln // 16 + 1, 16), dtype=np.uint8)
            rec[:, 0] = np.arange(rec.shape[0]) & 0xFF
            rec[:, 1] = (np.arange(rec.shape[0]) >> 8) & 0xFF
            rec[:, 4:8] = rng.integers(0, 4, (rec.shape[0], 4))
            b = rec.tobytes()[:ln]
        parts.append(b)
        n += len(b)
    return b"".join(parts)[:size]


def lzma1_filter(lc=3, lp=0, pb=2, dict_size=8 << 20, preset=6):
    return {"id": lzma.FILTER_LZMA1, "dict_size": dict_size, "lc": lc, "lp": lp, "pb": pb, "preset": preset}


def compress_alone(data: bytes, lc=3, lp=0, pb=2, dict_size=8 << 20, preset=6, size_mode: str = "eos") -> bytes:
    """.lzma stream via liblzma (always EOS marker + unknown size, Appendix B).
    size_mode: 'eos' (as written) or 'eos+size' (header size patched in)."""
    s = lzma.compress(data, format=lzma.FORMAT_ALONE, filters=[lzma1_filter(lc, lp, pb, dict_size, preset)])
    if size_mode == "eos+size":
        s = s[:5] + struct.pack("<Q", len(data)) + s[13:]
    elif size_mode != "eos":
        raise ValueError(size_mode)
    return s


def compress_raw_lzma2(data: bytes, lc=3, lp=0, pb=2, dict_size=8 << 20, preset=6) -> bytes:
    """Raw LZMA2 stream (chunk framing only; ends with 0x00)."""
    f = {"id": lzma.FILTER_LZMA2, "dict_size": dict_size, "lc": lc, "lp": lp, "pb": pb, "preset": preset}
    return lzma.compress(data, format=lzma.FORMAT_RAW, filters=[f])


def lzma2_with_resets(blocks, **kw) -> bytes:
    """One raw LZMA2 stream with a dictionary reset at each block: independently
    compressed raw-LZMA2 streams, each but the last stripped of its 0x00 (Appendix B)."""
    parts = [compress_raw_lzma2(b, **kw) for b in blocks]
    return b"".join(p[:-1] for p in parts[:-1]) + parts[-1]


def alone_from_lzma2_chunk(data: bytes, lc=3, lp=0, pb=2, dict_size=1 << 16, preset=6):
    """Known size, NO EOS marker: re-wrap the payload of a single-chunk raw-LZMA2
    stream as .lzma (Appendix B).  Returns None if the data does not fit one chunk."""
    s = compress_raw_lzma2(data, lc, lp, pb, dict_size, preset)
    c = s[0]
    if c < 0xE0:
        return None
    us = (((c & 0x1F) << 16) | (s[1] << 8) | s[2]) + 1
    cs = ((s[3] << 8) | s[4]) + 1
    if us != len(data) or 6 + cs + 1 != len(s) or s[-1] != 0:
        return None
    return bytes([s[5]]) + struct.pack("<I", dict_size) + struct.pack("<Q", us) + s[6:6 + cs]


# ---------- parallel builders for bench.py ----------

def _job_text_alone(args):
    import zlib
    seed, size, kw = args
    d = text_block(seed, size)
    return compress_alone(d, **kw), zlib.crc32(d)


def _job_text_lzma2(args):
    seed, size, kw = args
    return compress_raw_lzma2(text_block(seed, size), **kw)


def build_alone_streams(n: int, size: int, seed0: int = 0, workers: int | None = None, with_crc: bool = False, **kw):
    """n independent .lzma streams of `size` text-like bytes (stream i uses seed seed0+i).
    with_crc: also return the CRC32 of each plaintext."""
    workers = workers or os.cpu_count() or 1
    jobs = [(seed0 + i, size, kw) for i in range(n)]
    if workers == 1 or n < 4:
        out = [_job_text_alone(j) for j in jobs]
    else:
        with ProcessPoolExecutor(max_workers=min(workers, n)) as ex:
            out = list(ex.map(_job_text_alone, jobs, chunksize=max(1, n // (workers * 8))))
    streams, crcs = [o[0] for o in out], [o[1] for o in out]
    return (streams, crcs) if with_crc else streams


def build_lzma2_stream(n_blocks: int, block: int, seed0: int = 0, workers: int | None = None, **kw) -> bytes:
    """One raw LZMA2 stream of n_blocks x block text-like bytes, dict reset per block."""
    workers = workers or os.cpu_count() or 1
    jobs = [(seed0 + i, block, kw) for i in range(n_blocks)]
    if workers == 1 or n_blocks < 4:
        parts = [_job_text_lzma2(j) for j in jobs]
    else:
        with ProcessPoolExecutor(max_workers=min(workers, n_blocks)) as ex:
            parts = list(ex.map(_job_text_lzma2, jobs, chunksize=max(1, n_blocks // (workers * 8))))
    return b"".join(p[:-1] for p in parts[:-1]) + parts[-1]


# ---------- BASELINE config 5: many distinct streams of varied compressibility, cheap to generate ----------
# 1 024 x 4 MiB of text_block() alone costs ~8 core-minutes before any compression, so the plaintexts are cut
# from a per-process pool: 64 KiB slices at seeded byte offsets of 16 text blocks, a seed-dependent share of them
# replaced by incompressible noise or by mixed_block() material (runs, short periods, records).  Slices of one
# stream may overlap in the pool, which gives the far, long matches real archives have.  Compressed size per
# 4 MiB ranges over roughly 0.25 .. 0.5, so the size-balanced scheduler has something to balance.
_POOL = None
_POOL_BLOCKS, _POOL_BLOCK = 16, 2 << 20
_SLICE = 64 << 10


def _pool():
    global _POOL
    if _POOL is None:
        text = np.frombuffer(b"".join(text_block(50_000 + i, _POOL_BLOCK) for i in range(_POOL_BLOCKS)), dtype=np.uint8)
        mixed = np.frombuffer(mixed_block(60_000, 4 << 20), dtype=np.uint8)
        _POOL = (text, mixed)
    return _POOL


def varied_block(seed: int, size: int) -> bytes:
    text, mixed = _pool()
    rng = np.random.default_rng(11_000_027 * (seed + 1))
    p_noise = 0.03 * (seed % 8)             # 0 .. 21 % of the slices incompressible
    p_mixed = 0.04 * ((seed // 8) % 4)      # 0 .. 12 % runs / periods / records
    out = np.empty(size, dtype=np.uint8)
    pos = 0
    while pos < size:
        n = min(_SLICE, size - pos)
        r = rng.random()
        if r < p_noise:
            out[pos:pos + n] = rng.integers(0, 256, n, dtype=np.uint8)
        elif r < p_noise + p_mixed:
            o = int(rng.integers(0, mixed.size - n))
            out[pos:pos + n] = mixed[o:o + n]
        else:
            o = int(rng.integers(0, text.size - n))
            out[pos:pos + n] = text[o:o + n]
        pos += n
    return out.tobytes()


def _job_varied_alone(args):
    import zlib
    seed, size, kw = args
    d = varied_block(seed, size)
    return compress_alone(d, **kw), zlib.crc32(d)


def build_varied_streams(n: int, size: int, seed0: int = 0, workers: int | None = None, **kw):
    """n .lzma streams of varied_block(seed0 + i, size); returns (streams, plaintext CRC32s)."""
    workers = workers or os.cpu_count() or 1
    jobs = [(seed0 + i, size, kw) for i in range(n)]
    if workers == 1 or n < 4:
        out = [_job_varied_alone(j) for j in jobs]
    else:
        with ProcessPoolExecutor(max_workers=min(workers, n), initializer=_pool) as ex:
            out = list(ex.map(_job_varied_alone, jobs, chunksize=max(1, n // (workers * 8))))
    return [o[0] for o in out], [o[1] for o in out]


# ---------- BASELINE config 4: the mixed batch ----------
def mixed_batch(seed: int = 4, unit_size: int = 192 << 10, assets_dir: str | None = None):
    """>= 256 work items of every kind the decode path accepts, as a list of dicts:
        {"name", "kind": "alone" | "raw" | "lzma2", "data": bytes, "cap": int,  + raw: lc lp pb dict unpack | lzma2: dict}
    .lzma streams over the whole lc/lp/pb range liblzma writes (lc+lp <= 4) and, by re-labelling the property
    byte, beyond it (lc+lp up to 12: literal tables in HBM; such a stream decodes to garbage or an error -- parity
    with the oracle is what counts); EOS-only, size-only and EOS+size streams; headerless LZMA1 units (the sevenzip
    path); raw LZMA2 streams with dictionary resets whose incompressible blocks liblzma stores as uncompressed
    chunks; incompressible .lzma streams (9 bits per byte); the reference's bad_* assets, bit-flipped and truncated
    streams, bad headers."""
    import itertools
    import random
    rng = random.Random(seed)
    items = []
    k = 0
    goods = []
    for lc, lp, pb in itertools.product(range(5), range(5), range(5)):       # 75 combinations
        if lc + lp > 4:
            continue
        d = (text_block if k % 3 else mixed_block)(7000 + k, unit_size - 1000 * (k % 5))
        mode = ("eos", "eos+size")[k % 2]
        s = compress_alone(d, lc, lp, pb, 1 << (16 + k % 5), preset=(1, 6)[k % 2], size_mode=mode)
        items.append({"name": f"alone_lc{lc}lp{lp}pb{pb}_{mode}", "kind": "alone", "data": s, "cap": len(d) + (k % 3)})
        goods.append((s, len(d)))
        k += 1
    base = compress_alone(text_block(7100, 60_000), 4, 0, 2, 1 << 16)
    for lc, lp, pb in [(5, 0, 0), (8, 0, 2), (4, 4, 4), (8, 4, 4), (6, 2, 1), (5, 3, 3), (7, 1, 0), (4, 1, 2)]:
        items.append({"name": f"relabel_lc{lc}lp{lp}pb{pb}", "kind": "alone", "data": bytes([(pb * 5 + lp) * 9 + lc]) + base[1:], "cap": 40_000})
    for i, n in enumerate((1, 17, 300, 5000, 30_000, 60_000)):               # known size, no EOS marker
        s = alone_from_lzma2_chunk(text_block(7200 + i, n))
        if s:
            items.append({"name": f"size_no_eos_{n}", "kind": "alone", "data": s, "cap": n + (i % 2) * 100})
    for i in range(24):                                                       # headerless LZMA1 (sevenzip path)
        lc, lp, pb = [(3, 0, 2), (0, 0, 0), (2, 2, 4), (4, 0, 0)][i % 4]
        d = (text_block, mixed_block)[i % 2](7300 + i, unit_size + 333 * i)
        s = compress_alone(d, lc, lp, pb, 1 << 20, preset=1 + i % 6)
        unpack = len(d) if i % 3 else UNKNOWN_SIZE
        items.append({"name": f"raw_{i}", "kind": "raw", "data": s[13:], "cap": len(d), "lc": lc, "lp": lp, "pb": pb,
                      "dict": 1 << 20, "unpack": unpack})
    for i in range(24):                                                       # incompressible: 9 adaptive bits per byte
        d = random_block(7400 + i, unit_size // 2)
        items.append({"name": f"noise_{i}", "kind": "alone", "data": compress_alone(d, preset=1), "cap": len(d)})
    for i in range(24):                                                       # LZMA2 with resets and uncompressed chunks
        blocks = []
        for j in range(4):
            kind = (i + j) % 3
            blocks.append((text_block, random_block, mixed_block)[kind](7500 + 10 * i + j, unit_size // 2 + 4099 * j))
        lc, lp, pb = [(3, 0, 2), (0, 2, 0), (4, 0, 4), (1, 1, 1)][i % 4]
        s = lzma2_with_resets(blocks, lc=lc, lp=lp, pb=pb, dict_size=1 << 20, preset=1 + i % 6)
        items.append({"name": f"lzma2_{i}", "kind": "lzma2", "data": s, "cap": sum(map(len, blocks)), "dict": 1 << 20})
    s = lzma2_with_resets([text_block(7600, 100_000), random_block(7601, 90_000)], dict_size=1 << 20)
    items.append({"name": "lzma2_truncated", "kind": "lzma2", "data": s[:len(s) * 2 // 3], "cap": 190_000, "dict": 1 << 20})
    items.append({"name": "lzma2_no_terminator", "kind": "lzma2", "data": s[:-1], "cap": 190_000, "dict": 1 << 20})
    items.append({"name": "lzma2_q6_control", "kind": "lzma2", "data": s[:-1] + b"\x05rest", "cap": 190_000, "dict": 1 << 20})
    if assets_dir:
        for n in sorted(os.listdir(assets_dir)):
            if n.endswith(".lzma"):
                items.append({"name": "asset_" + n, "kind": "alone", "data": open(os.path.join(assets_dir, n), "rb").read(), "cap": 2 << 20})
            elif n.endswith(".lzma2"):
                items.append({"name": "asset_" + n, "kind": "lzma2", "data": open(os.path.join(assets_dir, n), "rb").read(), "cap": 2 << 20, "dict": 0})
    for i in range(60):                                                       # corrupt streams
        s, n = goods[i % len(goods)]
        b = bytearray(s)
        op = ("flip", "flip", "trunc", "byte", "size", "prop")[i % 6]
        if op == "flip":
            b[rng.randrange(13, len(b))] ^= 1 << rng.randrange(8)
        elif op == "trunc":
            del b[rng.randrange(13, len(b)):]
        elif op == "byte":
            b[rng.randrange(13, len(b))] = rng.randrange(256)
        elif op == "size":
            b[5:13] = struct.pack("<Q", max(0, n + rng.choice([-5000, -1, 1, 5000])))
        else:
            b[0] = rng.randrange(256)
        items.append({"name": f"corrupt_{i}_{op}", "kind": "alone", "data": bytes(b), "cap": n + 50_000})
    items.append({"name": "empty_input", "kind": "alone", "data": b"", "cap": 16})
    items.append({"name": "header_only", "kind": "alone", "data": goods[0][0][:13], "cap": 16})
    items.append({"name": "overflow", "kind": "alone", "data": goods[1][0], "cap": 1000})
    return items
