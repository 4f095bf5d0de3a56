"""ctypes wrapper of tests/enc (test-only greedy LZMA1 encoder)."""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None

EOS, SIZE, Q4_START, RAW = 1, 2, 4, 8


def lib():
    global _lib
    if _lib is None:
        subprocess.check_call(["make", "-C", os.path.join(_HERE, "enc"), "-s"])
        _lib = C.CDLL(os.path.join(_HERE, "enc", "_build", "liblzma_test_encoder.so"))
        _lib.lzma_test_encode.restype = C.c_size_t
        _lib.lzma_test_encode.argtypes = [C.c_char_p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_int,
                                          C.c_char_p, C.c_size_t]
    return _lib


def encode(data: bytes, lc=3, lp=0, pb=2, dict_size=1 << 16, flags=EOS) -> bytes:
    cap = len(data) * 2 + 4096
    out = C.create_string_buffer(cap)
    n = lib().lzma_test_encode(data, len(data), lc, lp, pb, dict_size, flags, out, cap)
    assert n > 0, "encoder overflow"
    return out.raw[:n]
