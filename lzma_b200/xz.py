"""`.xz` container front-end over the batch engine.

The reference package has no `.xz` reader; this is the caller that naturally produces many
independent units for the hot path: every block of an `.xz` file is a raw LZMA2 stream that starts
with a dictionary reset (`xz -T` / `--block-size` write one block per slice of the input), and the
index at the end of the file says where each block starts and how large it decodes -- so all
blocks of a file, and all files of a batch, go to the GPU in ONE `lzgpu_decode_batch` call.

Host side (this file): container parsing per the .xz file format 1.1.0 (stream header / footer,
index, block headers, LZMA2 filter properties), the CRC32 of the container fields, and SHA-256 block
checks.  Device side: the same units and kernels as everything else (`lzgpu_scan_lzma2` cuts each
block's payload at its dictionary resets), and the CRC-32 / CRC-64 of each block's decoded bytes
(`lzgpu_decode_batch_sums`: computed where the bytes lie, no host pass over the payload).
Only the LZMA2 filter alone is supported (what `xz` writes by default); BCJ / delta chains are
reported as unsupported.
"""
from __future__ import annotations

import ctypes as C
import ctypes.util
import hashlib
import struct
import zlib
from dataclasses import dataclass, field

import numpy as np

from . import _lib as L
from . import batch as B

MAGIC = b"\xfd7zXZ\x00"
FOOTER_MAGIC = b"YZ"
CHECK_NONE, CHECK_CRC32, CHECK_CRC64, CHECK_SHA256 = 0x0, 0x1, 0x4, 0xA
_CHECK_SIZE = {0: 0, 1: 4, 2: 4, 3: 4, 4: 8, 5: 8, 6: 8, 7: 16, 8: 16, 9: 16, 10: 32, 11: 32, 12: 32, 13: 64, 14: 64, 15: 64}
FILTER_LZMA2 = 0x21


class XZError(Exception):
    """Malformed or unsupported container (the compressed payload itself reports lzgpu statuses)."""


@dataclass
class Block:
    offset: int            # of the block header in the file
    header_size: int
    unpadded_size: int     # header + compressed data + check (from the index)
    uncompressed_size: int # from the index
    dict_size: int
    check_type: int
    data_off: int = 0      # compressed data = file[data_off : data_off + data_len]
    data_len: int = 0
    check: bytes = b""
    out_off: int = 0       # of the block's bytes in the decoded output of the whole file


@dataclass
class Stream:
    offset: int
    check_type: int
    blocks: list = field(default_factory=list)


def _vli(buf: bytes, pos: int, end: int):
    """Variable-length integer (at most 9 bytes, 63 bits)."""
    v = 0
    for i in range(9):
        if pos >= end:
            raise XZError("truncated variable-length integer")
        b = buf[pos]
        pos += 1
        v |= (b & 0x7F) << (7 * i)
        if not b & 0x80:
            if b == 0 and i > 0:
                raise XZError("non-minimal variable-length integer")
            return v, pos
    raise XZError("variable-length integer too long")


def lzma2_dict_size(props: int) -> int:
    """LZMA2 filter property byte -> dictionary size (same coding as reader2.go:296-298 for props <= 40)."""
    if props > 40:
        raise XZError("LZMA2 dictionary size property out of range")
    if props == 40:
        return 0xFFFFFFFF
    return (2 | (props & 1)) << (props // 2 + 11)


def _parse_block_header(data: bytes, off: int, check_type: int) -> tuple[int, int]:
    """Returns (header_size, dict_size)."""
    if off >= len(data) or data[off] == 0:
        raise XZError("block header expected")
    hs = (data[off] + 1) * 4
    if off + hs > len(data):
        raise XZError("truncated block header")
    hdr = data[off:off + hs]
    if zlib.crc32(hdr[:-4]) != struct.unpack("<I", hdr[-4:])[0]:
        raise XZError("block header CRC32 mismatch")
    flags = hdr[1]
    if flags & 0x3C:
        raise XZError("reserved block flags set")
    nfilters = (flags & 3) + 1
    pos, end = 2, hs - 4
    if flags & 0x40:
        _, pos = _vli(hdr, pos, end)
    if flags & 0x80:
        _, pos = _vli(hdr, pos, end)
    dict_size = None
    for k in range(nfilters):
        fid, pos = _vli(hdr, pos, end)
        psz, pos = _vli(hdr, pos, end)
        props = hdr[pos:pos + psz]
        pos += psz
        if pos > end:
            raise XZError("truncated filter properties")
        if fid != FILTER_LZMA2 or k != nfilters - 1 or nfilters != 1:
            raise XZError(f"unsupported filter chain (filter id {fid:#x}): only LZMA2 alone is supported")
        if psz != 1:
            raise XZError("LZMA2 filter properties must be one byte")
        dict_size = lzma2_dict_size(props[0])
    if any(hdr[pos:end]):
        raise XZError("non-zero block header padding")
    return hs, dict_size


def scan_xz(data: bytes) -> list[Stream]:
    """Locate every block of every stream of an .xz file from the footers and indexes (no decoding)."""
    streams: list[Stream] = []
    end = len(data)
    while end > 0:
        while end >= 4 and data[end - 4:end] == b"\0\0\0\0":   # stream padding (after any stream)
            end -= 4
        if end == 0:
            break
        if end < 32:
            raise XZError("file too short for an .xz stream")
        foot = data[end - 12:end]
        if foot[10:12] != FOOTER_MAGIC:
            raise XZError("stream footer magic missing")
        if zlib.crc32(foot[4:10]) != struct.unpack("<I", foot[0:4])[0]:
            raise XZError("stream footer CRC32 mismatch")
        index_size = (struct.unpack("<I", foot[4:8])[0] + 1) * 4
        flags = foot[8:10]
        if flags[0] != 0 or flags[1] & 0xF0:
            raise XZError("unsupported stream flags")
        check_type = flags[1] & 0x0F
        idx_off = end - 12 - index_size
        if idx_off < 12:
            raise XZError("index does not fit the file")
        idx = data[idx_off:idx_off + index_size]
        if idx[0] != 0:
            raise XZError("index indicator missing")
        if zlib.crc32(idx[:-4]) != struct.unpack("<I", idx[-4:])[0]:
            raise XZError("index CRC32 mismatch")
        nrec, pos = _vli(idx, 1, index_size - 4)
        recs = []
        for _ in range(nrec):
            unpadded, pos = _vli(idx, pos, index_size - 4)
            usize, pos = _vli(idx, pos, index_size - 4)
            if unpadded < 5:
                raise XZError("index record too small")
            recs.append((unpadded, usize))
        if any(idx[pos:index_size - 4]) or (pos + 3) // 4 * 4 != index_size - 4:
            raise XZError("bad index padding")
        blocks_size = sum((u + 3) // 4 * 4 for u, _ in recs)
        start = idx_off - blocks_size - 12
        if start < 0:
            raise XZError("blocks do not fit the file")
        head = data[start:start + 12]
        if head[:6] != MAGIC:
            raise XZError("stream header magic missing")
        if head[6:8] != flags or zlib.crc32(head[6:8]) != struct.unpack("<I", head[8:12])[0]:
            raise XZError("stream header does not match the footer")
        st = Stream(start, check_type)
        off = start + 12
        csz = _CHECK_SIZE[check_type]
        for unpadded, usize in recs:
            hs, dict_size = _parse_block_header(data, off, check_type)
            b = Block(off, hs, unpadded, usize, dict_size, check_type)
            b.data_off = off + hs
            b.data_len = unpadded - hs - csz
            if b.data_len < 0:
                raise XZError("index record smaller than its block header and check")
            pad = -(hs + b.data_len) % 4                 # block padding sits between the data and the check
            cpos = off + hs + b.data_len + pad
            if any(data[cpos - pad:cpos]):
                raise XZError("non-zero block padding")
            b.check = data[cpos:cpos + csz]
            st.blocks.append(b)
            off = cpos + csz
        streams.append(st)
        end = start
    streams.reverse()
    out = 0
    for st in streams:
        for b in st.blocks:
            b.out_off = out
            out += b.uncompressed_size
    return streams


_crc64_fn = None


def crc64(data: bytes) -> int:
    """CRC-64/XZ (ECMA-182 polynomial, reflected).  liblzma's routine when the shared library is
    there, else a table-driven fallback."""
    global _crc64_fn
    if _crc64_fn is None:
        try:
            lib = C.CDLL(ctypes.util.find_library("lzma") or "liblzma.so.5")
            lib.lzma_crc64.restype = C.c_uint64
            lib.lzma_crc64.argtypes = [C.c_char_p, C.c_size_t, C.c_uint64]
            _crc64_fn = lambda d: lib.lzma_crc64(d, len(d), 0)   # noqa: E731
        except (OSError, AttributeError):
            table = []
            for i in range(256):
                c = i
                for _ in range(8):
                    c = (c >> 1) ^ (0xC96C5795D7870F42 if c & 1 else 0)
                table.append(c)

            def slow(d, table=table):
                c = 0xFFFFFFFFFFFFFFFF
                for x in d:
                    c = table[(c ^ x) & 0xFF] ^ (c >> 8)
                return c ^ 0xFFFFFFFFFFFFFFFF
            _crc64_fn = slow
    return _crc64_fn(bytes(data))


def check_ok(check_type: int, expected: bytes, payload) -> bool:
    if check_type == CHECK_NONE:
        return True
    if check_type == CHECK_CRC32:
        return struct.pack("<I", zlib.crc32(payload)) == expected
    if check_type == CHECK_CRC64:
        return struct.pack("<Q", crc64(payload)) == expected
    if check_type == CHECK_SHA256:
        return hashlib.sha256(payload).digest() == expected
    raise XZError(f"unsupported integrity check id {check_type}")


def build_units(files: list[bytes]):
    """Units of every block of every file, over ONE input buffer (the files back to back, 16-byte
    aligned) and one output buffer.  Returns (units, in_buf, out_size, per_file) where per_file[i] =
    (streams, out_off, out_len, [(block, first_unit, n_units), ...])."""
    units, per_file = [], []
    in_off = out_off = 0
    chunks = []
    for data in files:
        streams = scan_xz(data)
        blocks = []
        f_out = out_off
        for st in streams:
            for b in st.blocks:
                payload = data[b.data_off:b.data_off + b.data_len]
                us, total, sst = B.scan_lzma2(payload, b.dict_size)
                if sst != L.OK or total != b.uncompressed_size:
                    raise XZError("block payload does not match its index record")
                for u in us:
                    u.in_off += in_off + b.data_off
                    u.out_off += f_out + b.out_off
                blocks.append((b, len(units), len(us)))
                units.extend(us)
        n_out = sum(b.uncompressed_size for st in streams for b in st.blocks)
        per_file.append((streams, f_out, n_out, blocks))
        chunks.append((in_off, data))
        in_off = (in_off + len(data) + 15) // 16 * 16
        out_off = (f_out + n_out + 15) // 16 * 16
    in_buf = np.zeros(max(in_off, 16), dtype=np.uint8)
    for off, data in chunks:
        in_buf[off:off + len(data)] = np.frombuffer(data, dtype=np.uint8)
    return units, in_buf, max(out_off, 16), per_file


def decode_xz_files(ctx, files: list[bytes], verify: bool = True) -> list[bytes]:
    """Decode a batch of .xz files: all blocks of all files in one call.  CRC-32 / CRC-64 block checks are
    computed by the GPU from the decoded bytes where they lie (lzgpu_decode_batch_sums: one checksum per
    unit, folded per block with lzgpu_crc32_combine / lzgpu_crc64_combine); only SHA-256 blocks are hashed
    on the host."""
    units, in_buf, out_size, per_file = build_units(files)
    out_buf = B._out_buffer(out_size)
    if verify:
        for _streams, _f_out, _n_out, blocks in per_file:
            for b, first, cnt in blocks:
                flag = {CHECK_CRC32: L.UF_SUM_CRC32, CHECK_CRC64: L.UF_SUM_CRC64}.get(b.check_type, 0)
                for k in range(first, first + cnt):
                    units[k].flags |= flag
    if hasattr(ctx, "decode_batch_sums"):
        res, _, sums = ctx.decode_batch_sums(units, in_buf, out_buf)
    else:                                   # the lane emulation of the CPU test tier has no checksum kernels
        res, _ = ctx.decode_batch(units, in_buf, out_buf)
        sums = None
    lib = L.lib()
    outs = []
    for fi, (streams, f_out, n_out, blocks) in enumerate(per_file):
        for b, first, cnt in blocks:
            got = 0
            for k in range(first, first + cnt):
                if res[k].status != L.OK:
                    raise XZError(f"file {fi}: block at {b.offset}: unit {k - first}: {L.status_name(res[k].status)} (site {res[k].err_site})")
                got += res[k].bytes_out
            if got != b.uncompressed_size:
                raise XZError(f"file {fi}: block at {b.offset}: decoded {got} bytes, index says {b.uncompressed_size}")
            if not verify:
                continue
            if sums is not None and b.check_type in (CHECK_CRC32, CHECK_CRC64):
                acc = 0
                for k in range(first, first + cnt):
                    if b.check_type == CHECK_CRC32:
                        acc = lib.lzgpu_crc32_combine(acc, int(sums[k]) & 0xFFFFFFFF, res[k].bytes_out)
                    else:
                        acc = lib.lzgpu_crc64_combine(acc, int(sums[k]), res[k].bytes_out)
                ok = acc == int.from_bytes(b.check, "little")
            else:
                ok = check_ok(b.check_type, b.check, out_buf[f_out + b.out_off:f_out + b.out_off + got])
            if not ok:
                raise XZError(f"file {fi}: block at {b.offset}: integrity check mismatch")
        outs.append(out_buf[f_out:f_out + n_out].tobytes())
    return outs


def decode_xz(ctx, data: bytes, verify: bool = True) -> bytes:
    return decode_xz_files(ctx, [data], verify)[0]
