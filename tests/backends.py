"""Two ways to run a batch of units in the tests:

* ``gpu``  -- the product: liblzgpu.so through the C ABI (lzma_b200.batch.Context);
* ``emu``  -- tests/emu: the kernel's unit code compiled for the host with lanes emulated
  (test infrastructure for the CPU-only tier; it is not shipped and not a fallback).
"""
import ctypes as C
import os

import numpy as np

from lzma_b200 import _lib as L
from lzma_b200._lib import Result, Unit

_EMU = None


def emu_lib():
    global _EMU
    if _EMU is None:
        here = os.path.dirname(os.path.abspath(__file__))
        _EMU = C.CDLL(os.path.join(here, "emu", "_build", "liblzgpu_emu.so"))
        _EMU.emu_decode_batch.restype = C.c_int
        _EMU.emu_decode_batch.argtypes = [C.POINTER(Unit), C.c_int64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64,
                                          C.POINTER(Result), C.c_int]
    return _EMU


class EmuContext:
    n_devices = 1

    def __init__(self, variant: int = 1):
        self.variant = variant

    def decode_batch(self, units, in_buf: np.ndarray, out_buf: np.ndarray):
        n = len(units)
        arr = units if isinstance(units, C.Array) else (Unit * max(n, 1))(*units)
        # the device path parses .lzma headers on the host inside lzgpu_decode_batch
        for i in range(n):
            if arr[i].kind == L.KIND_LZMA1_ALONE:
                u = arr[i]
                hdr = bytes(in_buf[u.in_off:u.in_off + min(u.in_len, 13)])
                st = L.lib().lzgpu_parse_alone_header(C.cast(C.c_char_p(hdr), C.c_void_p), len(hdr), C.byref(u))
                arr[i] = u
                if st != L.OK:
                    arr[i].user = 0xBAD00000 | st
        res = (Result * max(n, 1))()
        run = [i for i in range(n) if (arr[i].user >> 16) != 0xBAD0]
        sub = (Unit * max(len(run), 1))(*[arr[i] for i in run])
        sres = (Result * max(len(run), 1))()
        rc = emu_lib().emu_decode_batch(sub, len(run), in_buf.ctypes.data, in_buf.nbytes, out_buf.ctypes.data,
                                        out_buf.nbytes, sres, self.variant)
        assert rc == 0, rc
        for k, i in enumerate(run):
            res[i] = sres[k]
        for i in range(n):
            if (arr[i].user >> 16) == 0xBAD0:
                res[i].status = arr[i].user & 0xFFFF
        return res, None

    def close(self):
        pass


def make_context(kind: str, variant: int = 1):
    if kind == "emu":
        return EmuContext(variant)
    from lzma_b200.batch import Context
    return Context()
