"""Seeded test inputs shared by the CPU tier (oracle vs lane-emulated kernel code) and the GPU
tier (oracle vs liblzgpu.so).  Sizes are chosen so that the oracle finishes in seconds."""
import itertools
import os
import random
import struct

from lzma_b200 import corpus as K

ASSETS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_assets")


def asset(name: str) -> bytes:
    with open(os.path.join(ASSETS, name), "rb") as f:
        return f.read()


ALONE_ASSETS = ["a.lzma", "a_eos.lzma", "a_eos_and_size.lzma", "a_lp1_lc2_pb1.lzma", "bad_corrupted.lzma",
                "bad_eos_incorrect_size.lzma", "bad_incorrect_size.lzma", "randomfile.dat.lzma"]


def alone_cases(seed: int = 1, heavy: bool = True):
    """List of (name, stream, out_cap)."""
    rng = random.Random(seed)
    cases = [(n, asset(n), 2 << 20) for n in ALONE_ASSETS]
    blk = K.mixed_block(3, 50_000) + K.text_block(5, 30_000)
    good = []
    for lc, lp, pb in itertools.product(range(5), range(5), range(5)):
        if lc + lp > 4 or (not heavy and rng.random() < 0.7):
            continue
        mode = rng.choice(["eos", "eos+size"])
        s = K.compress_alone(blk, lc, lp, pb, 1 << 16, preset=rng.choice([1, 4, 6]), size_mode=mode)
        good.append(s)
        cases.append((f"props_lc{lc}_lp{lp}_pb{pb}_{mode}", s, len(blk) + 7))
    # dictionary smaller than the data: window wrap, isFull, far rep distances
    for ds in (4096, 8192, 1 << 16):
        d = K.mixed_block(11 + ds, 200_000)
        cases.append((f"dict_{ds}", K.compress_alone(d, 3, 0, 2, ds, preset=6), len(d)))
    # long runs: overlapping copies (dist < len), 273-byte matches
    runs = b"".join(bytes([i]) * (300 + 37 * i) for i in range(40)) + b"ab" * 5000 + b"xyz" * 3000 + bytes(100_000)
    cases.append(("runs", K.compress_alone(runs, preset=9), len(runs)))
    # known size, no EOS marker (re-wrapped single LZMA2 chunk)
    for i, n in enumerate((1, 2, 17, 5000, 60_000)):
        d = K.text_block(40 + i, n)
        s = K.alone_from_lzma2_chunk(d)
        if s:
            cases.append((f"size_no_eos_{n}", s, n))
            cases.append((f"size_no_eos_{n}_cap_large", s, n + 1000))
    # empty payloads / tiny streams
    cases.append(("empty_eos", K.compress_alone(b""), 16))
    cases.append(("one_byte", K.compress_alone(b"Z"), 16))
    cases.append(("zero_size_header", K.compress_alone(b"", size_mode="eos+size"), 16))
    # output capacity too small
    cases.append(("overflow_unknown_size", K.compress_alone(blk), 1000))
    cases.append(("overflow_known_size", K.compress_alone(blk, size_mode="eos+size"), 1000))
    # header-level failures
    cases.append(("no_bytes", b"", 16))
    cases.append(("bad_prop", bytes([225]) + bytes(12) + bytes(8), 16))
    cases.append(("bad_prop_only", bytes([255]), 16))
    cases.append(("short_header", good[0][:7], 16))
    cases.append(("header_only", good[0][:13], 16))
    cases.append(("rc_preamble_short", good[0][:16], 16))
    cases.append(("rc_first_byte_nonzero", good[0][:13] + b"\x01" + good[0][14:], 100_000))
    # size in header larger / smaller than the data
    s = K.compress_alone(blk, size_mode="eos+size")
    cases.append(("size_too_big", s[:5] + struct.pack("<Q", len(blk) + 5) + s[13:], len(blk) + 100))
    cases.append(("size_too_small", s[:5] + struct.pack("<Q", len(blk) - 5) + s[13:], len(blk) + 100))
    # truncations and bit flips of valid streams
    for k, s in enumerate(good[:10] if heavy else good[:4]):
        for r in range(3):
            cut = rng.randrange(14, len(s))
            cases.append((f"trunc_{k}_{r}", s[:cut], len(blk) + 300_000))
            b = bytearray(s)
            i = rng.randrange(13, len(b))
            b[i] ^= 1 << rng.randrange(8)
            cases.append((f"flip_{k}_{r}", bytes(b), len(blk) + 300_000))
            b = bytearray(s)
            i = rng.randrange(0, 13)
            b[i] ^= 1 << rng.randrange(8)
            cases.append((f"hdrflip_{k}_{r}", bytes(b), len(blk) + 300_000))
    return cases


def lzma2_cases(seed: int = 2):
    """List of (name, stream, dict_size, out_cap)."""
    rng = random.Random(seed)
    cases = [("asset_randomfile", asset("randomfile.dat.lzma2"), 0, 2 << 20)]
    blocks = [K.text_block(i, 300_000) for i in range(3)] + [K.random_block(1, 100_000), K.mixed_block(2, 200_000)]
    total = sum(map(len, blocks))
    s = K.lzma2_with_resets(blocks, dict_size=1 << 20)
    cases.append(("five_units", s, 1 << 20, total))
    cases.append(("five_units_small_dictarg", s, 0, total))
    one = K.compress_raw_lzma2(K.text_block(9, 700_000) + K.random_block(3, 70_000) + K.text_block(10, 100_000), dict_size=1 << 20)
    cases.append(("one_unit_many_chunks", one, 1 << 20, 900_000))
    for lc, lp, pb in ((0, 0, 0), (4, 0, 4), (0, 4, 2), (2, 2, 1), (1, 3, 3)):
        d = K.mixed_block(20 + lc, 150_000)
        cases.append((f"props_lc{lc}_lp{lp}_pb{pb}", K.compress_raw_lzma2(d, lc, lp, pb, 1 << 16, preset=5), 1 << 16, len(d)))
    small_dict = K.compress_raw_lzma2(K.mixed_block(77, 300_000), dict_size=4096)
    cases.append(("dict_4096", small_dict, 4096, 300_000))
    cases.append(("empty", K.compress_raw_lzma2(b""), 0, 16))
    cases.append(("nothing", b"", 0, 16))
    cases.append(("q6_control_0x03_ends_stream", s[:-1] + b"\x03garbage", 1 << 20, total))
    # truncation: inside a header, inside a payload, just before the terminator
    for r in range(6):
        cut = rng.randrange(1, len(s))
        cases.append((f"trunc_{r}", s[:cut], 1 << 20, total))
    cases.append(("trunc_no_terminator", s[:-1], 1 << 20, total))
    cases.append(("trunc_header_2", s[:2], 1 << 20, total))
    return cases


def encoder_cases(seed: int = 3, heavy: bool = True):
    """Streams only our test encoder can write (tests/enc): every prop byte the reference accepts
    (lc 0-8, lp 0-4, pb 0-4 -- liblzma stops at lc+lp = 4), size-only / EOS-only / both, odd
    dictionary sizes (Q3: wrapped-position contexts), a match at position 0 (Q4), long runs."""
    import encoder as E
    rng = random.Random(seed)
    d = K.mixed_block(5, 9_000) + K.text_block(6, 6_000)
    cases = []
    for lc, lp, pb in itertools.product(range(9), range(5), range(5)):
        if not heavy and (lc * 7 + lp * 3 + pb) % 5:
            continue
        fl = rng.choice([E.EOS, E.SIZE, E.EOS | E.SIZE])
        cases.append((f"enc_lc{lc}_lp{lp}_pb{pb}_f{fl}", E.encode(d, lc, lp, pb, 1 << 16, fl), len(d) + 3))
    for ds in (4096, 4097, 4100, 5000, 7777, 65_537):
        dd = K.mixed_block(9 + ds, 60_000)
        cases.append((f"enc_dict_{ds}", E.encode(dd, 3, 1, 2, ds, E.EOS), len(dd)))
        cases.append((f"enc_dict_{ds}_pb4_lp4", E.encode(dd, 0, 4, 4, ds, E.SIZE), len(dd)))
    z = bytes(300) + K.text_block(1, 2_000)
    cases.append(("enc_q4_match_at_position_0", E.encode(z, 3, 0, 2, 1 << 16, E.EOS | E.Q4_START), len(z)))
    z2 = bytes(3000)
    cases.append(("enc_q4_all_zero", E.encode(z2, 2, 2, 0, 4096, E.SIZE | E.Q4_START), len(z2)))
    runs = b"".join(bytes([i]) * (280 + 31 * i) for i in range(30)) + b"abcd" * 4000 + bytes(50_000)
    cases.append(("enc_runs", E.encode(runs, 3, 0, 2, 1 << 20, E.EOS | E.SIZE), len(runs)))
    big = K.text_block(77, 400_000)
    cases.append(("enc_text_400k_dict64k", E.encode(big, 3, 0, 2, 1 << 16, E.EOS), len(big)))
    s = E.encode(d, 8, 4, 4, 1 << 16, E.EOS | E.SIZE)
    for r in range(4 if heavy else 1):   # corrupt streams with huge literal tables
        b = bytearray(s)
        i = rng.randrange(13, len(b))
        b[i] ^= 1 << rng.randrange(8)
        cases.append((f"enc_flip_lc8lp4_{r}", bytes(b), len(d) + 50_000))
        cases.append((f"enc_trunc_lc8lp4_{r}", s[:rng.randrange(14, len(s))], len(d) + 50_000))
    return cases


def fuzz_cases(n: int = 1200, seed: int = 7):
    """Hostile variants of valid .lzma streams long enough for the fast decoder to be running when the damage
    is met: single / multiple bit flips, byte overwrites, inserted and deleted bytes, truncations, wrong sizes,
    wrong dictionary sizes, swapped property bytes.  List of (name, stream, out_cap)."""
    rng = random.Random(seed)
    bases = []
    for i, (lc, lp, pb, ds, kind, size) in enumerate([
            (3, 0, 2, 1 << 20, "text", 160_000), (0, 2, 0, 1 << 16, "mixed", 120_000), (4, 0, 4, 4096, "text", 90_000),
            (1, 3, 1, 1 << 18, "mixed", 200_000), (3, 0, 2, 1 << 16, "random", 20_000), (2, 1, 3, 8192, "runs", 150_000)]):
        if kind == "text":
            d = K.text_block(300 + i, size)
        elif kind == "mixed":
            d = K.mixed_block(300 + i, size)
        elif kind == "random":
            d = K.random_block(300 + i, size)
        else:
            d = (b"".join(bytes([j]) * (50 + 13 * j) for j in range(60)) + b"abcd" * 9000 + K.text_block(9, 40_000))[:size]
        mode = "eos" if i % 2 else "eos+size"
        bases.append((K.compress_alone(d, lc, lp, pb, ds, preset=rng.choice([1, 6]), size_mode=mode), len(d)))
    cases = []
    for k in range(n):
        s, size = bases[k % len(bases)]
        b = bytearray(s)
        op = rng.choice(["flip", "flip", "flip", "flips", "byte", "insert", "delete", "trunc", "size", "dict", "prop"])
        if op == "flip":
            b[rng.randrange(13, len(b))] ^= 1 << rng.randrange(8)
        elif op == "flips":
            for _ in range(rng.randrange(2, 6)):
                b[rng.randrange(13, len(b))] ^= 1 << rng.randrange(8)
        elif op == "byte":
            b[rng.randrange(13, len(b))] = rng.randrange(256)
        elif op == "insert":
            b.insert(rng.randrange(18, len(b)), rng.randrange(256))
        elif op == "delete":
            del b[rng.randrange(18, len(b))]
        elif op == "trunc":
            del b[rng.randrange(13, len(b)):]
        elif op == "size":
            b[5:13] = struct.pack("<Q", max(0, size + rng.choice([-70_000, -300, -1, 1, 2, 300, 70_000])))
        elif op == "dict":
            b[1:5] = struct.pack("<I", rng.choice([0, 1, 4096, 4097, 5000, 65_535, 1 << 20]))
        else:
            b[0] = rng.randrange(225)
        cases.append((f"fuzz_{k}_{op}", bytes(b), size + rng.choice([0, 1, 4096, 300_000])))
    return cases
