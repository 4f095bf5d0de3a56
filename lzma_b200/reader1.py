"""Host-side mirror of reader1.go: the .lzma (LZMA1 "alone") reader API over the GPU batch engine.

Same names, argument meaning and error behaviour as the reference (reader1.go:10-254):
constructors read the header and the range-coder preamble eagerly and report their errors;
body errors surface from Read.  Where the reference decodes symbol by symbol as Read is
called, this reader hands the whole stream to the CUDA library on the first Read and then
serves from the decoded buffer; there is no CPU decode path."""
from __future__ import annotations

import numpy as np

from . import _lib as L
from . import batch as B
from . import errors as E
from .readcloser import readCloser

_default_ctx = None


def default_context():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = B.Context()
    return _default_ctx


def _status_error(status: int):
    """lzgpu status -> the error value the reference returns."""
    if status in (L.OK, L.OK_INPUT_EXHAUSTED):
        return None                       # OK_INPUT_EXHAUSTED: clean EOF in the reference (Q1)
    if status == L.RESULT_ERROR:
        return E.ErrResultError
    if status == L.INCORRECT_PROPERTIES:
        return E.ErrIncorrectProperties
    if status == L.UNEXPECTED_EOF:
        return E.ErrUnexpectedEOF
    if status == L.OUTPUT_OVERFLOW:
        return E.ErrOutputOverflow
    return E.Error(f"lzgpu: unexpected status {status}")


def DecodeProp(d: int):
    """reader1.go:210-221.  Returns (lc, pb, lp, err) -- the reference's order."""
    if d >= 9 * 5 * 5:
        return 0, 0, 0, E.ErrIncorrectProperties
    lc = d % 9
    d //= 9
    return lc, d // 5, d % 5, None


def DecodeDictSize(properties: bytes):
    """reader1.go:193-208."""
    d = int.from_bytes(bytes(properties[:4]), "little")
    return max(d, 1 << 12), None


def DecodeUnpackSize(header: bytes) -> int:
    """reader1.go:178-191."""
    return int.from_bytes(bytes(header[:8]), "little")


def _read_exact(stream, n: int) -> bytes:
    out = b""
    while len(out) < n:
        c = stream.read(n - len(out))
        if not c:
            break
        out += c
    return out


class Reader1:
    """reader1.go:10-16."""

    def __init__(self, ctx=None):
        self._ctx = ctx
        self._in = None
        self._lc = self._lp = self._pb = 0
        self._dict = 0
        self._unpack = L.UNKNOWN_SIZE
        self._preamble = b""
        self._out = None          # decoded bytes, once the GPU has run
        self._pos = 0
        self._err = None          # error to report once the decoded bytes are served
        self.isEndOfStream = False

    # -- Reader1.initialize (reader1.go:149-159): range-coder preamble --
    def _initialize(self):
        p = _read_exact(self._in, 5)
        if len(p) < 5:
            if len(p) >= 1 and p[0] != 0:
                return E.Errorf("rangeDec.Init", E.ErrResultError)
            return E.Errorf("rangeDec.Init", E.EOF)
        if p[0] != 0:
            return E.Errorf("rangeDec.Init", E.ErrResultError)   # range_decoder.go:32-34
        self._preamble = p
        return None

    def _decode(self):
        payload = self._preamble + self._in.read()
        ctx = self._ctx or default_context()
        u = L.Unit()
        u.kind = L.KIND_LZMA1_RAW
        u.lc, u.lp, u.pb = self._lc, self._lp, self._pb
        u.dict_size = self._dict
        u.unpack_size = self._unpack
        u.in_off, u.in_len = 0, len(payload)
        in_buf = np.frombuffer(payload, dtype=np.uint8) if payload else np.zeros(1, dtype=np.uint8)
        # The header's size field is untrusted (13 bytes can claim 2^50): the first capacity is bounded by the
        # payload and grows towards the declared size only when the decoder asks for more.  A retry decodes from the
        # start again; as each attempt is 8x the previous one, all failed attempts together cost < 1/7 of the last.
        known = self._unpack != L.UNKNOWN_SIZE
        cap = max(1 << 16, 8 * len(payload))
        if known:
            cap = min(cap, max(self._unpack, 1))
        while True:
            u.out_off, u.out_cap = 0, cap
            out = B._out_buffer(max(cap, 16))
            res, _ = ctx.decode_batch([u], in_buf, out)
            r = res[0]
            can_grow = cap < self._unpack if known else cap < (1 << 40)
            if r.status == L.OUTPUT_OVERFLOW and can_grow:
                cap = min(cap * 8, self._unpack) if known else cap * 8   # the streaming reader has no capacity: grow, decode again
                continue
            break
        self._out = out[:r.bytes_out]
        self._err = _status_error(r.status)

    def Read(self, p) -> tuple:
        """reader1.go:223-254.  p: a writable buffer (bytearray / memoryview / numpy uint8)."""
        if self._out is None:
            self._decode()
        n = min(len(p), len(self._out) - self._pos)
        if n:
            p[:n] = self._out[self._pos:self._pos + n].tobytes()
            self._pos += n
        if n == len(p) and n > 0:
            return n, None
        if self._err is not None:
            err, self._err = self._err, None
            self.isEndOfStream = True
            return n, err
        self.isEndOfStream = True
        return n, E.EOF

    # reader1.go:161-176 exist for Reader2's use of a shared Reader1; the GPU engine walks LZMA2
    # chunks on the device, so here they only re-arm the object.
    def Reset(self):
        self.isEndOfStream = False

    def Reopen(self, inStream, unpackSize: int):
        self.isEndOfStream = False
        self._in, self._unpack, self._out, self._pos, self._err = inStream, unpackSize, None, 0, None
        return self._initialize()


def NewReader1(inStream, ctx=None):
    """reader1.go:18-24 + initializeFull (:77-101).  inStream: a binary file-like object."""
    r = Reader1(ctx)
    r._in = inStream
    b = _read_exact(inStream, 1)
    if len(b) < 1:
        return r, E.EOF
    lc, pb, lp, err = DecodeProp(b[0])
    if err is not None:
        return r, E.Errorf("decode prop", err)
    ds = _read_exact(inStream, 4)
    if len(ds) < 4:
        return r, E.Errorf("decode dict size", E.EOF)
    us = _read_exact(inStream, 8)
    if len(us) < 8:
        return r, E.Errorf("decode unpack size", E.EOF)
    r._lc, r._lp, r._pb = lc, lp, pb
    r._dict, _ = DecodeDictSize(ds)
    r._unpack = DecodeUnpackSize(us)
    return r, r._initialize()


def NewLZMADecompressorForSevenZip(props: bytes, unpackSize: int, readers: list, ctx=None):
    """reader1.go:32-61: decompressor constructor with the bodgit/sevenzip signature."""
    if len(readers) != 1:
        return None, E.errNeedOneReader
    lc, pb, lp, err = DecodeProp(props[0])
    if err is not None:
        return None, err
    dict_size, err = DecodeDictSize(props[1:5])
    if err is not None:
        return None, err
    r = Reader1(ctx)
    r._in = readers[0]
    r._lc, r._lp, r._pb, r._dict, r._unpack = lc, lp, pb, dict_size, unpackSize
    return readCloser(readers[0], r), r._initialize()
