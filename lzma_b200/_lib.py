"""ctypes binding of liblzgpu.so (the C ABI declared in include/lzgpu.h).

The shared library is built in-tree (lzma_b200/liblzgpu.so) by
``__graft_entry__.build()`` / ``make -C lzma_b200/csrc``.  If it is missing the
import fails loudly: there is no Python or CPU decode path.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LZGPU_LIB") or os.path.join(_HERE, "liblzgpu.so")   # LZGPU_LIB: A/B builds of the same library (scripts/)

# enum lzgpu_status
OK, OK_INPUT_EXHAUSTED, RESULT_ERROR, INCORRECT_PROPERTIES, UNEXPECTED_EOF, OUTPUT_OVERFLOW = range(6)
DICT_OUT_OF_RANGE, UNEXPECTED_LZMA2_CODE = 6, 7
NOT_RUN = 255
# enum lzgpu_error
E_OK, E_NO_DEVICE, E_CUDA, E_INVALID, E_NOMEM = 0, -1, -2, -3, -4
# enum lzgpu_kind
KIND_LZMA1_ALONE, KIND_LZMA1_RAW, KIND_LZMA2_GROUP = 0, 1, 2
UNKNOWN_SIZE = (1 << 64) - 1
UF_LZMA2_LAST, UF_LZMA2_FRESH, UF_BITS_KNOWN, UF_SUM_CRC32, UF_SUM_CRC64 = 1, 2, 4, 8, 16


class Unit(C.Structure):
    _fields_ = [("in_off", C.c_uint64), ("in_len", C.c_uint64), ("out_off", C.c_uint64), ("out_cap", C.c_uint64),
                ("unpack_size", C.c_uint64), ("dict_size", C.c_uint32), ("kind", C.c_uint8), ("lc", C.c_uint8),
                ("lp", C.c_uint8), ("pb", C.c_uint8), ("lit_bits", C.c_uint8), ("pos_bits", C.c_uint8), ("pad8", C.c_uint8 * 2),
                ("flags", C.c_uint32), ("user", C.c_uint64)]


class Result(C.Structure):
    _fields_ = [("status", C.c_int32), ("err_site", C.c_int32), ("bytes_out", C.c_uint64), ("bytes_in", C.c_uint64),
                ("final_code", C.c_uint32), ("device", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("kernel_ms", C.c_double), ("h2d_ms", C.c_double), ("d2h_ms", C.c_double), ("total_ms", C.c_double),
                ("bytes_in", C.c_uint64), ("bytes_out", C.c_uint64), ("launches", C.c_int32), ("devices", C.c_int32)]


assert C.sizeof(Unit) == 64 and C.sizeof(Result) == 32

# every symbol include/lzgpu.h declares: (name, restype, argtypes)
_vp, _u8p, _u64, _i64 = C.c_void_p, C.c_void_p, C.c_uint64, C.c_int64
SYMBOLS = [
    ("lzgpu_abi_version", C.c_int, []),
    ("lzgpu_device_count", C.c_int, []),
    ("lzgpu_status_name", C.c_char_p, [C.c_int]),
    ("lzgpu_last_error", C.c_char_p, []),
    ("lzgpu_decode_prop", C.c_int, [C.c_uint8, C.POINTER(C.c_uint8), C.POINTER(C.c_uint8), C.POINTER(C.c_uint8)]),
    ("lzgpu_decode_dict_size", C.c_uint32, [C.c_char_p]),
    ("lzgpu_decode_unpack_size", C.c_uint64, [C.c_char_p]),
    ("lzgpu_decode_dict_size2", C.c_uint32, [C.c_uint8]),
    ("lzgpu_parse_alone_header", C.c_int, [_u8p, _u64, C.POINTER(Unit)]),
    ("lzgpu_scan_lzma2", _i64, [_u8p, _u64, C.c_uint32, C.POINTER(Unit), _i64, C.POINTER(C.c_uint64), C.POINTER(C.c_int32)]),
    ("lzgpu_shard_units", C.c_int, [C.POINTER(Unit), _i64, C.c_int, C.POINTER(C.c_int32)]),
    ("lzgpu_ctx_create", C.c_int, [C.POINTER(C.c_int), C.c_int, C.POINTER(_vp)]),
    ("lzgpu_ctx_destroy", None, [_vp]),
    ("lzgpu_ctx_device_count", C.c_int, [_vp]),
    ("lzgpu_decode_batch", C.c_int, [_vp, C.POINTER(Unit), _i64, _u8p, _u64, _u8p, _u64, C.POINTER(Result), C.POINTER(Stats)]),
    ("lzgpu_plan_create", C.c_int, [_vp, C.c_int, C.POINTER(Unit), _i64, _u64, _u64, C.POINTER(_vp)]),
    ("lzgpu_plan_launch", C.c_int, [_vp, _u8p, _u8p, _vp]),
    ("lzgpu_plan_results", C.c_int, [_vp, C.POINTER(Result), C.POINTER(Stats)]),
    ("lzgpu_plan_launch_count", C.c_int, [_vp]),
    ("lzgpu_plan_crc32", C.c_int, [_vp, _u8p, C.POINTER(C.c_uint32)]),
    ("lzgpu_plan_crc64", C.c_int, [_vp, _u8p, C.POINTER(C.c_uint64)]),
    ("lzgpu_decode_batch_sums", C.c_int, [_vp, C.POINTER(Unit), _i64, _u8p, _u64, _u8p, _u64, C.POINTER(Result), C.POINTER(Stats), C.POINTER(C.c_uint64)]),
    ("lzgpu_crc32_combine", C.c_uint32, [C.c_uint32, C.c_uint32, _u64]),
    ("lzgpu_crc64_combine", C.c_uint64, [C.c_uint64, C.c_uint64, _u64]),
    ("lzgpu_alloc_pinned", C.c_void_p, [_u64]),
    ("lzgpu_free_pinned", None, [_vp]),
    ("lzgpu_plan_destroy", None, [_vp]),
]

_lib = None


class LzgpuError(RuntimeError):
    """Infrastructure failure of the CUDA library (not a per-unit decode outcome)."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"lzgpu error {code}: {msg}")
        self.code = code


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build the CUDA extension first "
                "(python -c 'import __graft_entry__ as g; g.build()' or make -C lzma_b200/csrc). "
                "lzma_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, res, args in SYMBOLS:
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        if L.lzgpu_abi_version() != 2:
            raise ImportError("liblzgpu.so ABI version mismatch")
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != E_OK:
        raise LzgpuError(rc, lib().lzgpu_last_error().decode("utf-8", "replace"))


def status_name(s: int) -> str:
    return lib().lzgpu_status_name(int(s)).decode()
