"""Batch entry point: many independent units -> one call into the CUDA library.

Host-side mirror of what a Go caller does through cgo: build
``lzgpu_unit`` descriptors, hand two flat buffers to ``lzgpu_decode_batch``, read
per-unit ``lzgpu_result``s.  All decoding happens in liblzgpu.so on the GPU.
"""
from __future__ import annotations

import ctypes as C
import threading
from dataclasses import dataclass
from typing import Iterable, Sequence

import numpy as np

from . import _lib as L
from ._lib import Result, Stats, Unit


@dataclass
class UnitResult:
    status: int
    err_site: int
    data: bytes
    bytes_in: int
    final_code: int
    device: int

    @property
    def status_name(self) -> str:
        return L.status_name(self.status)

    @property
    def ok(self) -> bool:
        return self.status in (L.OK, L.OK_INPUT_EXHAUSTED)


def pinned_empty(n: int) -> np.ndarray:
    """uint8 array in page-locked host memory (lzgpu_alloc_pinned): buffers of this kind take the batch call's
    zero-copy / streamed route.  Falls back to ordinary memory when pinning fails (the call still works)."""
    import weakref
    n = max(int(n), 1)
    p = L.lib().lzgpu_alloc_pinned(n)
    if not p:
        return np.empty(n, dtype=np.uint8)
    carr = (C.c_uint8 * n).from_address(p)
    weakref.finalize(carr, L.lib().lzgpu_free_pinned, p)
    return np.frombuffer(carr, dtype=np.uint8)


def _out_buffer(n: int) -> np.ndarray:
    return pinned_empty(n) if n >= (32 << 20) else np.empty(max(n, 16), dtype=np.uint8)


def _ptr(a: np.ndarray) -> int:
    return a.ctypes.data


def parse_alone_header(data, unit: Unit | None = None):
    """Reader1.initializeFull's header parse (reader1.go:77-101).  Returns (status, unit)."""
    u = unit or Unit()
    buf = bytes(data[:13])
    st = L.lib().lzgpu_parse_alone_header(C.cast(C.c_char_p(buf), C.c_void_p), len(buf), C.byref(u))
    return st, u


def scan_lzma2(data: bytes, dict_size: int = 0):
    """Host chunk scanner (reader2.go:100-214): returns (units, total_out, stream_status)."""
    lib = L.lib()
    total, sst = C.c_uint64(), C.c_int32()
    buf = C.cast(C.c_char_p(data), C.c_void_p)
    n = lib.lzgpu_scan_lzma2(buf, len(data), dict_size, None, 0, C.byref(total), C.byref(sst))
    if n < 0:
        L.check(int(n))
    arr = (Unit * max(int(n), 1))()
    n2 = lib.lzgpu_scan_lzma2(buf, len(data), dict_size, arr, n, C.byref(total), C.byref(sst))
    assert n2 == n
    return [arr[i] for i in range(n)], total.value, sst.value


def shard_units(units: Sequence[Unit], n_shards: int) -> list[int]:
    """LPT assignment by compressed size; identical on every rank."""
    n = len(units)
    arr = (Unit * max(n, 1))(*units)
    out = (C.c_int32 * max(n, 1))()
    L.check(L.lib().lzgpu_shard_units(arr, n, n_shards, out))
    return [out[i] for i in range(n)]


class Context:
    """Owns streams and staging buffers on the chosen GPUs (lzgpu_ctx)."""

    def __init__(self, devices: Iterable[int] | None = None):
        lib = L.lib()
        self._lock = threading.Lock()     # one call at a time per context; close() waits for a call in flight on another thread
        self._h = C.c_void_p()
        if devices is None:
            rc = lib.lzgpu_ctx_create(None, 0, C.byref(self._h))
        else:
            d = list(devices)
            arr = (C.c_int * len(d))(*d)
            rc = lib.lzgpu_ctx_create(arr, len(d), C.byref(self._h))
        L.check(rc)

    @property
    def n_devices(self) -> int:
        return L.lib().lzgpu_ctx_device_count(self._h)

    def close(self):
        with self._lock:
            if self._h:
                L.lib().lzgpu_ctx_destroy(self._h)
                self._h = C.c_void_p()

    def _handle(self):
        if not self._h:
            raise L.LzgpuError(L.E_INVALID, "the context has been closed")
        return self._h

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- host buffers ----
    def decode_batch(self, units: Sequence[Unit], in_buf: np.ndarray, out_buf: np.ndarray):
        """lzgpu_decode_batch over numpy uint8 buffers (pinned or pageable).  Returns (results, stats)."""
        n = len(units)
        arr = units if isinstance(units, C.Array) else (Unit * max(n, 1))(*units)
        res = (Result * max(n, 1))()
        st = Stats()
        with self._lock:
            L.check(L.lib().lzgpu_decode_batch(self._handle(), arr, n, _ptr(in_buf), in_buf.nbytes, _ptr(out_buf), out_buf.nbytes, res, C.byref(st)))
        return res, st

    def decode_batch_sums(self, units: Sequence[Unit], in_buf: np.ndarray, out_buf: np.ndarray):
        """lzgpu_decode_batch_sums: as decode_batch, plus the CRC-32 / CRC-64 of every unit flagged UF_SUM_CRC32 /
        UF_SUM_CRC64, computed on the GPU.  Returns (results, stats, sums[n] as numpy uint64)."""
        n = len(units)
        arr = units if isinstance(units, C.Array) else (Unit * max(n, 1))(*units)
        res = (Result * max(n, 1))()
        st = Stats()
        sums = (C.c_uint64 * max(n, 1))()
        with self._lock:
            L.check(L.lib().lzgpu_decode_batch_sums(self._handle(), arr, n, _ptr(in_buf), in_buf.nbytes, _ptr(out_buf), out_buf.nbytes, res, C.byref(st), sums))
        return res, st, np.frombuffer(sums, dtype=np.uint64)[:n].copy()

    # ---- device buffers ----
    def plan(self, units: Sequence[Unit], in_size: int, out_size: int, dev_index: int = 0) -> "Plan":
        return Plan(self, units, in_size, out_size, dev_index)


class Plan:
    """lzgpu_plan: descriptors uploaded once, launch as often as needed."""

    def __init__(self, ctx: Context, units: Sequence[Unit], in_size: int, out_size: int, dev_index: int = 0):
        self.ctx = ctx
        self.n = len(units)
        arr = units if isinstance(units, C.Array) else (Unit * max(self.n, 1))(*units)
        self._h = C.c_void_p()
        L.check(L.lib().lzgpu_plan_create(ctx._h, dev_index, arr, self.n, in_size, out_size, C.byref(self._h)))

    @property
    def launch_count(self) -> int:
        return L.lib().lzgpu_plan_launch_count(self._h)

    def launch(self, d_in_ptr: int, d_out_ptr: int, stream: int = 0):
        L.check(L.lib().lzgpu_plan_launch(self._h, d_in_ptr, d_out_ptr, stream))

    def results(self):
        res = (Result * max(self.n, 1))()
        st = Stats()
        L.check(L.lib().lzgpu_plan_results(self._h, res, C.byref(st)))
        return res, st

    def crc32(self, d_out_ptr: int) -> np.ndarray:
        """CRC-32 (zlib's) of every unit's decoded bytes, computed on the device (lzgpu_plan_crc32)."""
        crc = (C.c_uint32 * max(self.n, 1))()
        L.check(L.lib().lzgpu_plan_crc32(self._h, d_out_ptr, crc))
        return np.frombuffer(crc, dtype=np.uint32)[:self.n].copy()

    def crc64(self, d_out_ptr: int) -> np.ndarray:
        """CRC-64/XZ of every unit's decoded bytes, computed on the device (lzgpu_plan_crc64)."""
        crc = (C.c_uint64 * max(self.n, 1))()
        L.check(L.lib().lzgpu_plan_crc64(self._h, d_out_ptr, crc))
        return np.frombuffer(crc, dtype=np.uint64)[:self.n].copy()

    def close(self):
        if self._h:
            L.lib().lzgpu_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---------------------------------------------------------------- convenience layer

def _round_up(x: int, a: int) -> int:
    return (x + a - 1) // a * a


def build_alone_batch(streams: Sequence[bytes], out_caps: Sequence[int] | None = None, default_ratio: int = 8):
    """Lay .lzma streams into one input buffer (16-byte aligned slots) and size the output.

    Returns (units, in_buf, out_size, header_status).  A stream whose header is bad
    keeps its header status and gets a zero-capacity unit that is still submitted
    (the library reports the same status)."""
    units, hstat = [], []
    in_off = out_off = 0
    offs = []
    for i, s in enumerate(streams):
        st, u = parse_alone_header(s)
        u.kind = L.KIND_LZMA1_ALONE
        u.in_off, u.in_len = in_off, len(s)
        if out_caps is not None:
            cap = out_caps[i]
        elif st == L.OK and u.unpack_size != L.UNKNOWN_SIZE:
            cap = min(u.unpack_size, 1 << 40)
        else:
            cap = max(1 << 16, default_ratio * len(s))
        u.out_off, u.out_cap = out_off, cap
        units.append(u)
        hstat.append(st)
        offs.append(in_off)
        in_off = _round_up(in_off + len(s), 16)
        out_off = _round_up(out_off + cap, 16)
    in_buf = np.zeros(max(in_off, 16), dtype=np.uint8)
    for off, s in zip(offs, streams):
        in_buf[off:off + len(s)] = np.frombuffer(s, dtype=np.uint8)
    return units, in_buf, max(out_off, 16), hstat


def decode_alone_streams(ctx: Context, streams: Sequence[bytes], out_caps: Sequence[int] | None = None,
                         max_retries: int = 6) -> list[UnitResult]:
    """Decode independent .lzma streams; streams of unknown size that overflow their
    first capacity guess are retried with a larger one."""
    results: list[UnitResult | None] = [None] * len(streams)
    todo = list(range(len(streams)))
    caps = list(out_caps) if out_caps is not None else None
    ratio = 8
    for _ in range(max_retries + 1):
        sub = [streams[i] for i in todo]
        units, in_buf, out_size, _ = build_alone_batch(sub, [caps[i] for i in todo] if caps else None, ratio)
        out_buf = _out_buffer(out_size)
        res, _st = ctx.decode_batch(units, in_buf, out_buf)
        nxt = []
        for k, i in enumerate(todo):
            r, u = res[k], units[k]
            data = out_buf[u.out_off:u.out_off + r.bytes_out].tobytes()
            results[i] = UnitResult(r.status, r.err_site, data, r.bytes_in, r.final_code, r.device)
            if r.status == L.OUTPUT_OVERFLOW and caps is None and u.unpack_size == L.UNKNOWN_SIZE:
                nxt.append(i)
        if not nxt:
            break
        todo = nxt
        ratio *= 8
    return results  # type: ignore[return-value]


def decode_lzma2_stream(ctx: Context, data: bytes, dict_size: int = 0, as_array: bool = False, out_pool=None):
    """Decode one raw LZMA2 stream: scan into units, decode them in parallel,
    return (status, err_site, bytes).  The decoded bytes of units before the first
    failing one are returned with the failure, like the reference's reader would
    have delivered them."""
    units, total, sst = scan_lzma2(data, dict_size)
    in_buf = np.frombuffer(data, dtype=np.uint8) if len(data) else np.zeros(1, dtype=np.uint8)
    out_buf = out_pool(max(total, 16)) if out_pool else _out_buffer(max(total, 16))   # out_pool: the caller recycles (page-locked) buffers
    res, _ = ctx.decode_batch(units, in_buf, out_buf)
    n_out = 0
    fin = (lambda a: a) if as_array else (lambda a: a.tobytes())   # as_array: a view of the (possibly page-locked) buffer
    for u, r in zip(units, res):
        if r.status not in (L.OK,):
            n_out = u.out_off + r.bytes_out
            return r.status, r.err_site, fin(out_buf[:n_out])
        n_out = u.out_off + r.bytes_out
    return L.OK, 0, fin(out_buf[:n_out])
