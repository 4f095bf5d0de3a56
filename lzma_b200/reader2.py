"""Host-side mirror of reader2.go: the raw-LZMA2 reader API over the GPU batch engine.

NewReader2 reads the first chunk header eagerly like the reference (reader2.go:26-41,77-98).
Read then works in WAVES: the input is read incrementally, chunk headers are
walked on the host (reader2.go:100-214) until enough units -- chunk runs that begin at a dictionary
reset -- are buffered (`wave_bytes` of decoded output), that wave is decoded in parallel on the
GPU(s) and served; while a wave is being served the NEXT one is already being read and decoded on a
second thread (decode-ahead: the library call releases the GIL), so a steady reader sees the GPU's
throughput rather than decode and delivery taking turns.  Memory is bounded by two waves (unless the
stream never resets its dictionary), and a caller that stops reading early pays for at most one
wave beyond what it read."""
from __future__ import annotations

import threading

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
        self._buf = None
        self._rd = 0
        self._in_eof = False
        self._last = False
        self.wave_bytes = 1 << 30            # decoded bytes per GPU call (at least one unit): a wave takes as long as its longest unit, so large waves are what gives throughput
        self.decode_ahead = True             # decode wave k+1 on a second thread while wave k is served
        self._ahead = None                   # (thread, result box) of the wave being decoded ahead
        self._pool = []                      # output buffers of delivered waves (page-locked memory is dear to allocate)
        self._cur_buf = None

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

    # ---- incremental input ----
    def _fill(self, need: int) -> bool:
        """Make at least `need` unread bytes available in self._buf (False: the input ended first)."""
        while len(self._buf) - self._rd < need and not self._in_eof:
            c = self._in.read(max(1 << 20, need - (len(self._buf) - self._rd)))
            if not c:
                self._in_eof = True
                break
            self._buf += c
        return len(self._buf) - self._rd >= need

    def _independent_from(self, pos: int) -> bool:
        """Uncompressed chunk with dictionary reset at `pos`: does the first LZMA chunk after it (if any comes
        before the next reset) reset the state and carry properties?  Otherwise it inherits the previous
        unit's coder (reader2.go:155-165) and must stay in the same wave."""
        buf, p0 = self._buf, pos
        while True:
            self._rd = p0                                       # _fill counts from the wave's read position
            if not self._fill(pos - p0 + 3):
                return True
            c = buf[pos]
            if c == 0 or 3 <= c < 0x80 or c >= 0xE0 or (c == 1 and pos != p0):
                return True
            if c >= 0x80:
                return c >= 0xC0
            pos += 3 + ((buf[pos + 1] << 8) | buf[pos + 2]) + 1

    def _next_wave(self):
        """Walk chunk headers from the current position until `wave_bytes` of output are covered and
        the next chunk starts a new unit (dictionary reset).  Returns (bytes of the wave, last)."""
        buf = self._buf
        start = pos = self._rd
        out = 0
        first = True
        while True:
            self._rd = pos
            if not self._fill(1):
                return bytes(buf[start:]), True                 # ran off the input: the scanner reports it
            ctrl = buf[pos]
            if ctrl == 0 or 3 <= ctrl < 0x80:                   # end of stream (0x03-0x7F too: Q6)
                self._rd = pos + 1
                return bytes(buf[start:pos + 1]), True
            # a unit that inherits nothing: dictionary reset AND the first LZMA chunk after it brings new properties
            reset = ctrl >= 0xE0 or (ctrl == 1 and not first and out >= self.wave_bytes and self._independent_from(pos))
            if reset and not first and out >= self.wave_bytes:
                return bytes(buf[start:pos]) + b"\0", False     # wave ends before this chunk; terminate it
            hl = 3 if ctrl < 0x80 else (5 if ctrl < 0xC0 else 6)
            if not self._fill(hl):
                self._rd = len(buf)
                return bytes(buf[start:]), True
            usz = ((buf[pos + 1] << 8) | buf[pos + 2]) + 1
            if ctrl >= 0x80:
                usz += (ctrl & 0x1F) << 16
                payload = ((buf[pos + 3] << 8) | buf[pos + 4]) + 1
            else:
                payload = usz
            if not self._fill(hl + payload):
                self._rd = len(buf)
                return bytes(buf[start:]), True
            pos += hl + payload
            out += usz
            first = False

    def _decode_wave(self):
        """Reads the next wave's input and decodes it; returns (out, err, last).  Runs on the ahead thread too:
        only one _decode_wave is ever in flight, and nothing else touches self._in / self._buf meanwhile."""
        if self._buf is None:                                   # first wave: header bytes read by NewReader2
            self._buf = bytearray(self._head)
            self._rd = 0
            self._in_eof = False
        data, last = self._next_wave()
        if self._rd > (8 << 20):                                # drop what has been handed to the GPU
            del self._buf[:self._rd]
            self._rd = 0
        ctx = self._ctx or default_context()
        held = []

        def pool(n):
            held.append(self._buffer(n))
            return held[0]
        st, _site, out = B.decode_lzma2_stream(ctx, data, self._dict, as_array=True, out_pool=pool)
        err = _status_error(st)
        return out, err, last or err is not None, (held[0] if held else None)

    def _buffer(self, n: int):
        """An output buffer of at least n bytes: a delivered wave's if one fits (three rotate: served, decoding, spare)."""
        for k, b in enumerate(self._pool):
            if b.nbytes >= n:
                return self._pool.pop(k)
        self._pool.clear()
        return B._out_buffer(n)

    def close(self):
        """Waits for a wave being decoded ahead (call before closing the context the reader uses)."""
        if self._ahead is not None:
            self._ahead[0].join()
            self._ahead = None

    def _start_ahead(self):
        if not self.decode_ahead:
            return
        box = []

        def run():
            try:
                box.append(self._decode_wave())
            except BaseException as e:      # surfaces from the Read that needs this wave
                box.append(e)
        t = threading.Thread(target=run, daemon=True)
        t.start()
        self._ahead = (t, box)

    def _take_wave(self):
        if self._ahead is not None:
            t, box = self._ahead
            self._ahead = None
            t.join()
            if isinstance(box[0], BaseException):
                raise box[0]
            w = box[0]
        else:
            w = self._decode_wave()
        if self._cur_buf is not None:
            self._pool.append(self._cur_buf)                   # the delivered wave's buffer: reuse it
        self._out, self._err, self._last, self._cur_buf = w
        self._pos = 0
        if not self._last:
            self._start_ahead()

    def Read(self, p) -> tuple:
        """reader2.go:216-250."""
        if self._out is None:
            self._take_wave()
        while self._pos == len(self._out) and not self._last and len(p):
            self._take_wave()                                   # previous wave delivered: the one decoded meanwhile
        n = min(len(p), len(self._out) - self._pos)
        if n:
            if isinstance(p, np.ndarray):
                p[:n] = self._out[self._pos:self._pos + n]
            else:
                p[:n] = self._out[self._pos:self._pos + n].tobytes()
            self._pos += n
        if n == len(p) and n > 0:
            return n, None
        if not self._last:
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
