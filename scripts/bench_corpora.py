#!/usr/bin/env python
"""Device-timed throughput of the decode kernel on other shapes than bench.py's headline workload:
batch size (units per GPU), unit size, and data kind.  Inputs resident in HBM, CUDA events around
the launches, every unit's CRC checked.  Prints one JSON line per shape."""
import ctypes as C
import json
import os
import sys
import time
import zlib

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lzma_b200 import _lib as L  # noqa: E402
from lzma_b200 import batch as B  # noqa: E402
from lzma_b200 import corpus as K  # noqa: E402

KINDS = {"text": K.text_block, "random": K.random_block, "mixed": K.mixed_block}


def run(ctx, kind, n_units, size, distinct=32, steps=3):
    plains = [KINDS[kind](1000 + i, size) for i in range(distinct)]
    streams = [K.compress_alone(p) for p in plains]
    pick = [i % distinct for i in range(n_units)]
    units, in_buf, out_size, _ = B.build_alone_batch([streams[i] for i in pick], [size] * n_units)
    d_in = torch.from_numpy(in_buf).cuda()
    d_out = torch.empty(out_size, dtype=torch.uint8, device="cuda")
    plan = ctx.plan(units, in_buf.nbytes, out_size)
    st = torch.cuda.current_stream().cuda_stream or 1
    for _ in range(2):
        plan.launch(d_in.data_ptr(), d_out.data_ptr(), st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        plan.launch(d_in.data_ptr(), d_out.data_ptr(), st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    res, _ = plan.results()
    out = d_out.cpu().numpy()
    crcs = [zlib.crc32(p) for p in plains]
    for k in range(0, n_units, max(1, n_units // 64)):
        u = units[k]
        assert res[k].status == L.OK and zlib.crc32(out[u.out_off:u.out_off + size]) == crcs[pick[k]]
    comp = sum(len(streams[i]) for i in pick)
    plan.close()
    del d_in, d_out
    print(json.dumps({"kind": kind, "units": n_units, "unit_bytes": size, "ratio": round(n_units * size / comp, 2),
                      "ms": round(ms, 2), "GBps": round(n_units * size / ms / 1e6, 3)}), flush=True)


def run_replicated(ctx, n_units, size, distinct=16, steps=2):
    """BASELINE config 5 on ONE GPU (16 384 streams x 4 MiB = 64 GiB of output): `distinct` compressed streams are
    resident once and every unit decodes one of them into its OWN output range, so the kernel does the full work
    while the host builds only `distinct` streams.  A sample of units is copied back and CRC-checked."""
    plains = [K.text_block(3000 + i, size) for i in range(distinct)]
    streams = [K.compress_alone(p) for p in plains]
    t_units, in_buf, _, _ = B.build_alone_batch(streams, [size] * distinct)
    units = (L.Unit * n_units)()
    for k in range(n_units):
        t = t_units[k % distinct]
        u = units[k]
        C.memmove(C.byref(u), C.byref(t), C.sizeof(L.Unit))
        u.out_off = k * size
        u.out_cap = size
    out_size = n_units * size + 16
    d_in = torch.from_numpy(in_buf).cuda()
    d_out = torch.empty(out_size, dtype=torch.uint8, device="cuda")
    plan = ctx.plan(units, in_buf.nbytes, out_size)
    st = torch.cuda.current_stream().cuda_stream or 1
    plan.launch(d_in.data_ptr(), d_out.data_ptr(), st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        plan.launch(d_in.data_ptr(), d_out.data_ptr(), st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    res, _ = plan.results()
    bad = sum(1 for k in range(n_units) if res[k].status != L.OK or res[k].bytes_out != size)
    assert bad == 0, bad
    crcs = [zlib.crc32(p) for p in plains]
    # EVERY unit verified where it lies: CRC-32 on the device (lzgpu_plan_crc32) against zlib's of the plaintext;
    # a sample is also copied back and checked on the host
    t0 = time.perf_counter()
    dev_crc = plan.crc32(d_out.data_ptr())
    crc_ms = (time.perf_counter() - t0) * 1e3
    wrong = [k for k in range(n_units) if int(dev_crc[k]) != crcs[k % distinct]]
    assert not wrong, wrong[:8]
    for k in range(0, n_units, max(1, n_units // 96)):
        assert zlib.crc32(d_out[k * size:(k + 1) * size].cpu().numpy()) == crcs[k % distinct], k
    comp = sum(len(streams[k % distinct]) for k in range(n_units))
    plan.close()
    del d_in, d_out
    print(json.dumps({"kind": "text-replicated", "units": n_units, "unit_bytes": size, "distinct": distinct,
                      "ratio": round(n_units * size / comp, 2), "ms": round(ms, 2), "GBps": round(n_units * size / ms / 1e6, 3),
                      "verified": f"device CRC-32 of all {n_units} units ({crc_ms:.1f} ms = {n_units * size / crc_ms / 1e6:.0f} GB/s incl. launch + D2H of the CRCs)"}), flush=True)


def run_lzma2(ctx, n_blocks, size, distinct=32, steps=3):
    """BASELINE config 3: ONE raw LZMA2 stream with a dictionary reset every `size` bytes; the host scanner
    cuts it into units (chunk runs starting at a dictionary reset) that decode in parallel."""
    blocks = [K.text_block(2000 + i, size) for i in range(distinct)]
    parts = [K.compress_raw_lzma2(b) for b in blocks]
    seq = [i % distinct for i in range(n_blocks)]
    stream = b"".join(parts[i][:-1] for i in seq[:-1]) + parts[seq[-1]]
    units, total, sst = B.scan_lzma2(stream, 8 << 20)
    assert len(units) == n_blocks and total == n_blocks * size and sst == L.OK
    in_buf = np.frombuffer(stream + bytes(16), dtype=np.uint8)
    d_in = torch.from_numpy(in_buf.copy()).cuda()
    d_out = torch.empty(total + 16, dtype=torch.uint8, device="cuda")
    plan = ctx.plan(units, in_buf.nbytes, total + 16)
    st = torch.cuda.current_stream().cuda_stream or 1
    for _ in range(2):
        plan.launch(d_in.data_ptr(), d_out.data_ptr(), st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        plan.launch(d_in.data_ptr(), d_out.data_ptr(), st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    res, _ = plan.results()
    out = d_out.cpu().numpy()
    crcs = [zlib.crc32(b) for b in blocks]
    for k in range(0, n_blocks, max(1, n_blocks // 64)):
        assert res[k].status == L.OK and zlib.crc32(out[k * size:(k + 1) * size]) == crcs[seq[k]]
    plan.close()
    del d_in, d_out
    print(json.dumps({"kind": "lzma2-stream", "units": n_blocks, "unit_bytes": size, "ratio": round(total / len(stream), 2),
                      "ms": round(ms, 2), "GBps": round(total / ms / 1e6, 3)}), flush=True)


def run_uncompressed(ctx, n_units, size, steps=5):
    """An LZMA2 batch made of uncompressed chunks only (what xz writes for incompressible data; the reference's own
    benchmark of this path: reader2_test.go:31-36): every unit is a run of 64 KiB `0x01/0x02` chunks -- a copy, done by
    the warp's 16-byte mover (lzgpu_unit.cuh warp_copy_in)."""
    rng = np.random.default_rng(1)
    data = rng.integers(0, 256, size, dtype=np.uint8).tobytes()
    chunks = bytearray()
    for pos in range(0, size, 1 << 16):
        n = min(1 << 16, size - pos)
        chunks += bytes([1 if pos == 0 else 2, (n - 1) >> 8, (n - 1) & 0xFF]) + data[pos:pos + n]
    unit_stream = bytes(chunks)
    stream = unit_stream * n_units + b"\0"
    units, total, sst = B.scan_lzma2(stream, 8 << 20)
    assert len(units) == n_units and total == n_units * size and sst == L.OK
    in_buf = np.frombuffer(stream + bytes(16), dtype=np.uint8)
    d_in = torch.from_numpy(in_buf.copy()).cuda()
    d_out = torch.empty(total + 16, dtype=torch.uint8, device="cuda")
    plan = ctx.plan(units, in_buf.nbytes, total + 16)
    st = torch.cuda.current_stream().cuda_stream or 1
    for _ in range(3):
        plan.launch(d_in.data_ptr(), d_out.data_ptr(), st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        plan.launch(d_in.data_ptr(), d_out.data_ptr(), st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    res, _ = plan.results()
    assert all(res[k].status == L.OK and res[k].bytes_out == size for k in range(n_units))
    crc = plan.crc32(d_out.data_ptr())
    assert (crc == zlib.crc32(data)).all()
    plan.close()
    del d_in, d_out
    print(json.dumps({"kind": "lzma2-uncompressed-chunks", "units": n_units, "unit_bytes": size, "ms": round(ms, 3),
                      "GBps": round(total / ms / 1e6, 1), "copy_GBps_read_plus_write": round(2 * total / ms / 1e6, 1),
                      "verified": "device CRC-32 of every unit"}), flush=True)


if __name__ == "__main__":
    with B.Context([0]) as ctx:
        if "--config5" in sys.argv:        # 16 384 x 4 MiB on one GPU (64 GiB of output in HBM)
            run_replicated(ctx, 16384, 4 << 20)
            sys.exit(0)
        if "--shapes" in sys.argv:         # e.g. --shapes text:148,text:1024:4,random:148   (kind:units[:MiB per unit])
            for spec in sys.argv[sys.argv.index("--shapes") + 1].split(","):
                f = spec.split(":")
                run(ctx, f[0], int(f[1]), (int(f[2]) if len(f) > 2 else 1) << 20)
            sys.exit(0)
        if "--lzma2" in sys.argv:          # BASELINE config 3's shape, and twice as many units (14 per SM)
            run_lzma2(ctx, 1024, 1 << 20)
            run_lzma2(ctx, 2072, 1 << 20)
            sys.exit(0)
        if "--quick-random" in sys.argv:   # incompressible data: 9 adaptive bits per byte, literals only
            run(ctx, "random", 148, 1 << 20)
            run(ctx, "mixed", 1024, 1 << 20)
            sys.exit(0)
        if "--quick" in sys.argv:          # A/B of tuning variants: lone-warp latency and the bench shape
            for n in (148, 1024):
                run(ctx, "text", n, 1 << 20)
            sys.exit(0)
        if "--uncompressed" in sys.argv:
            run_uncompressed(ctx, 1024, 1 << 20)
            run_uncompressed(ctx, 4096, 1 << 20)
            sys.exit(0)
        for n in (148, 592, 1024, 1924, 2048, 2072, 4096, 8192):
            run(ctx, "text", n, 1 << 20)
        run(ctx, "text", 2048, 4 << 20, distinct=16)      # BASELINE config 5's per-GPU shape at 8 GPUs
        run(ctx, "random", 148, 1 << 20)
        run(ctx, "random", 1024, 1 << 20)
        run(ctx, "mixed", 1024, 1 << 20)
        run_lzma2(ctx, 1024, 1 << 20)                     # BASELINE config 3: one 1 GiB LZMA2 stream
        run_uncompressed(ctx, 1024, 1 << 20)
        run_uncompressed(ctx, 4096, 1 << 20)
