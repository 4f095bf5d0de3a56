"""Host-side mirror of reader2.go: the raw-LZMA2 reader API over the GPU batch engine.

NewReader2 reads the first chunk header eagerly like the reference (reader2.go:26-41,77-98);
the first Read scans the whole stream into units (chunk runs that begin at a dictionary
reset) and decodes them in parallel on the GPU(s)."""
from __future__ import annotations

import numpy as np

from . import _lib as L
from . import batch as B
from . import errors as E
from .reader1 import _read_exact, _status_error, default_context
from .readcloser import readCloser


def DecodeDictSize2(encodedDictSize: int) -> int:
    """reader2.go:296-298."""
    return ((2 | (encodedDictSize & 1)) << (encodedDictSize // 2 + 11)) & 0xFFFFFFFF


class Reader2:
    def __init__(self, inStream, dictSize: int, ctx=None):
        self._in = inStream
        self._dict = dictSize & 0xFFFFFFFF
        self._ctx = ctx
        self._head = b""
        self._out = None
        self._pos = 0
        self._err = None

    def _initialize(self):
        """validateDictSize + startChunk (reader2.go:77-173) as far as the first header."""
        if self._dict < (1 << 12):
            self._dict = 8 << 20
        c = _read_exact(self._in, 1)
        if len(c) < 1:
            return E.ErrUnexpectedEOF                        # reader2.go:103-110
        self._head = c
        ctrl = c[0]
        if ctrl == 0 or 3 <= ctrl < 0x80:
            return None
        hl = 3 if ctrl < 0x80 else (5 if ctrl < 0xC0 else 6)
        rest = _read_exact(self._in, hl - 1)
        self._head += rest
        if len(rest) < hl - 1:
            return E.ErrUnexpectedEOF                        # reader2.go:121-128
        if ctrl >= 0x80:
            # first LZMA chunk: NewReader1ForReader2 -> DecodeProp + rangeDec.Init (reader2.go:146-153)
            prop = self._head[5] if hl == 6 else 0
            if prop >= 225:
                return E.ErrIncorrectProperties
            pre = _read_exact(self._in, 1)
            self._head += pre
            if len(pre) < 1:
                return E.Errorf("rangeDec.Init", E.EOF)
            if pre[0] != 0:
                return E.Errorf("rangeDec.Init", E.ErrResultError)
        return None

    def _decode(self):
        data = self._head + self._in.read()
        ctx = self._ctx or default_context()
        st, _site, out = B.decode_lzma2_stream(ctx, data, self._dict)
        self._out = np.frombuffer(out, dtype=np.uint8)
        self._err = _status_error(st)

    def Read(self, p) -> tuple:
        """reader2.go:216-250."""
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
            return n, err
        return n, E.EOF


def NewReader2(inStream, dictSize: int, ctx=None):
    """reader2.go:26-41."""
    r = Reader2(inStream, dictSize, ctx)
    return r, r._initialize()


def NewLZMA2DecompressorForSevenZip(props: bytes, _unused: int, readers: list, ctx=None):
    """reader2.go:49-75."""
    if len(readers) != 1:
        return None, E.errNeedOneReader
    if len(props) != 1:
        return None, E.errInsufficientProperties
    r = Reader2(readers[0], DecodeDictSize2(props[0]), ctx)
    return readCloser(readers[0], r), r._initialize()
