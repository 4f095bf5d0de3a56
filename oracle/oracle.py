"""ctypes wrapper around oracle/_build/liblzma_oracle.so -- TEST INFRASTRUCTURE ONLY.

The oracle is a CPU restatement of kulaginds/lzma's decode path (see
lzma_oracle.c).  It may be imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / ``--impl reference`` legs, and by nothing under
lzma_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liblzma_oracle.so")

OK, OK_INPUT_EXHAUSTED, RESULT_ERROR, INCORRECT_PROPERTIES, UNEXPECTED_EOF, OUTPUT_OVERFLOW = range(6)
STATUS_NAMES = ["OK", "OK_INPUT_EXHAUSTED", "RESULT_ERROR", "INCORRECT_PROPERTIES", "UNEXPECTED_EOF", "OUTPUT_OVERFLOW"]
UNKNOWN_SIZE = (1 << 64) - 1


class _Result(C.Structure):
    _fields_ = [("status", C.c_int32), ("err_site", C.c_int32), ("bytes_out", C.c_uint64),
                ("bytes_in", C.c_uint64), ("final_code", C.c_uint32), ("pad", C.c_uint32)]


@dataclass
class Result:
    status: int
    err_site: int
    data: bytes
    bytes_in: int
    final_code: int

    @property
    def status_name(self) -> str:
        return STATUS_NAMES[self.status]


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (seconds).  Building the checker is not using it."""
    src = os.path.join(_HERE, "lzma_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(
            os.path.getmtime(src), os.path.getmtime(os.path.join(_HERE, "lzma_oracle.h"))):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        u8p, u64 = C.c_char_p, C.c_uint64
        L.orc_lzma_alone.argtypes = [u8p, u64, C.c_void_p, u64, C.POINTER(_Result)]
        L.orc_lzma_raw.argtypes = [u8p, u64, C.c_uint8, C.c_uint8, C.c_uint8, C.c_uint32, u64,
                                   C.c_void_p, u64, C.POINTER(_Result)]
        L.orc_lzma2.argtypes = [u8p, u64, C.c_uint32, C.c_void_p, u64, C.POINTER(_Result)]
        L.orc_decode_prop.argtypes = [C.c_uint8] + [C.POINTER(C.c_uint8)] * 3
        L.orc_decode_dict_size.argtypes = [u8p]
        L.orc_decode_dict_size.restype = C.c_uint32
        L.orc_decode_unpack_size.argtypes = [u8p]
        L.orc_decode_unpack_size.restype = C.c_uint64
        L.orc_decode_dict_size2.argtypes = [C.c_uint8]
        L.orc_decode_dict_size2.restype = C.c_uint32
        L.orc_lzma_alone_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        _lib = L
    return _lib


def _finish(res: _Result, buf) -> Result:
    n = min(res.bytes_out, len(buf))
    return Result(res.status, res.err_site, C.string_at(buf, n), res.bytes_in, res.final_code)


def lzma_alone(data: bytes, out_cap: int) -> Result:
    """NewReader1 + io.Copy on a .lzma stream (reader1.go:18, :223)."""
    buf = C.create_string_buffer(max(out_cap, 1))
    res = _Result()
    if lib().orc_lzma_alone(data, len(data), buf, out_cap, C.byref(res)) != 0:
        raise MemoryError("oracle allocation failed")
    return _finish(res, buf)


def lzma_raw(data: bytes, lc: int, lp: int, pb: int, dict_size: int, unpack_size: int, out_cap: int) -> Result:
    """NewLZMADecompressorForSevenZip path (reader1.go:32-61): headerless LZMA1."""
    buf = C.create_string_buffer(max(out_cap, 1))
    res = _Result()
    if lib().orc_lzma_raw(data, len(data), lc, lp, pb, dict_size, unpack_size, buf, out_cap, C.byref(res)) != 0:
        raise MemoryError("oracle allocation failed")
    return _finish(res, buf)


def lzma2(data: bytes, dict_size: int, out_cap: int) -> Result:
    """NewReader2 + io.Copy on a raw LZMA2 stream (reader2.go:26, :216)."""
    buf = C.create_string_buffer(max(out_cap, 1))
    res = _Result()
    if lib().orc_lzma2(data, len(data), dict_size, buf, out_cap, C.byref(res)) != 0:
        raise MemoryError("oracle allocation failed")
    return _finish(res, buf)


def decode_prop(d: int):
    """DecodeProp (reader1.go:210-221): returns (lc, pb, lp) in the reference's order."""
    lc, pb, lp = C.c_uint8(), C.c_uint8(), C.c_uint8()
    st = lib().orc_decode_prop(d, C.byref(lc), C.byref(pb), C.byref(lp))
    if st != OK:
        return None
    return lc.value, pb.value, lp.value


def decode_dict_size(b: bytes) -> int:
    return lib().orc_decode_dict_size(bytes(b[:4]))


def decode_unpack_size(b: bytes) -> int:
    return lib().orc_decode_unpack_size(bytes(b[:8]))


def decode_dict_size2(b: int) -> int:
    return lib().orc_decode_dict_size2(b)


def lzma_alone_batch(in_base, in_off, in_len, out_base, out_off, out_cap, threads: int):
    """Threaded batch decode over numpy arrays (bench baseline).  Returns (n_bad, statuses)."""
    import numpy as np
    n = len(in_off)
    res = (_Result * n)()
    io_ = np.ascontiguousarray(in_off, dtype=np.uint64)
    il = np.ascontiguousarray(in_len, dtype=np.uint64)
    oo = np.ascontiguousarray(out_off, dtype=np.uint64)
    oc = np.ascontiguousarray(out_cap, dtype=np.uint64)
    bad = lib().orc_lzma_alone_batch(in_base.ctypes.data, io_.ctypes.data, il.ctypes.data,
                                     out_base.ctypes.data, oo.ctypes.data, oc.ctypes.data,
                                     C.cast(res, C.c_void_p), n, threads)
    return bad, [(r.status, r.bytes_out) for r in res]
