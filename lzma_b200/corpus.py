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
