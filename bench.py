#!/usr/bin/env python
"""bench.py -- decompressed GB/s of the batch LZMA / LZMA2 decode path on B200.

Headline workload (BASELINE.json configs[1]): a batch of 1024 independent .lzma streams x 1 MiB of synthetic
text-like data, lc3 lp0 pb2, 8 MiB dictionary, compressed by liblzma preset 6, on one GPU.  With N GPUs (one
process per GPU under torchrun) the global batch is N x 1024 units, sharded over ranks by compressed size
(lzgpu_shard_units); there is no data-path collective (weak scaling).

A step = one pass of the decode path over the rank's units:
  value  : inputs and outputs resident in HBM, CUDA events around the launches
  e2e    : the same batch through lzgpu_decode_batch with pinned HOST buffers, H2D of the compressed input and D2H
           of the decoded output inside the timed region
  cpu_baseline / --impl reference : the C restatement of the reference's decoder (oracle/, kind "port": the Go
           toolchain is absent so the reference itself cannot run) on the box's host cores, bounded sample.

The same JSON line carries the other BASELINE configurations as sub-records, measured in the same run:
  config3 : ONE raw LZMA2 stream of 1 GiB with a dictionary reset every 1 MiB -> host scan -> 1 024 units,
            sharded over the ranks (device-timed, and end to end through scan + lzgpu_decode_batch)
  config4 : the mixed batch (every lc/lp/pb, EOS / size / both, headerless LZMA1, LZMA2 with uncompressed chunks,
            incompressible data, corrupt streams): status, error site and bytes of every unit against the oracle
  config5 : a FIXED global batch of 16 384 streams x 4 MiB (1 024 distinct seeds of varied compressibility, each
            decoded into 16 separate output ranges), LPT-sharded by compressed size over the N ranks, every unit
            verified by the on-device CRC-32: the strong-scaling point of `--gpus N`.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

STREAM_SIZE = 1 << 20
METRIC = "decompressed GB/s (batch, device-timed)"
# name of the dominant kernel as ncu prints it (variant 33 | V_PB2 = 97: the V_CHAIN decoder with compact posState
# tables, under the SM-resident scheduler), and the tag profiles/traffic.json must carry for its dram-bytes figure to
# be quoted
KERNEL_NAME = "void lzgpu_sm_kernel<97>(KArgs, SmArgs)"
KERNEL_TAG = "r02b"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ----------------------------------------------------------------------------- corpora (cached on local disk)
def _cache_dir() -> str:
    return os.environ.get("LZMA_B200_CACHE", "/tmp")


def _save(path, **arrs):
    try:
        tmp = path + f".{os.getpid()}.tmp.npz"
        np.savez(tmp, **arrs)
        os.replace(tmp, path)
    except OSError:
        pass


def _pack(streams):
    lens = np.array([len(s) for s in streams], dtype=np.int64)
    offs = np.zeros(len(streams), dtype=np.int64)
    np.cumsum(lens[:-1], out=offs[1:])
    return np.frombuffer(b"".join(streams), dtype=np.uint8), offs, lens


def build_corpus(distinct: int, size: int):
    """`distinct` .lzma streams (stream i: text_block(seed=i)), liblzma preset 6."""
    path = os.path.join(_cache_dir(), f"lzma_b200_corpus_text_{distinct}x{size}_lc3lp0pb2_d8M_p6.npz")
    if os.path.exists(path):
        z = np.load(path)
        return z["blob"], z["offs"], z["lens"], z["crc"]
    from lzma_b200 import corpus as K
    t0 = time.time()
    streams, crcs = K.build_alone_streams(distinct, size, seed0=0, with_crc=True)
    blob, offs, lens = _pack(streams)
    crc = np.array(crcs, dtype=np.uint32)
    log(f"[bench] config 2 corpus: {distinct} streams x {size} B in {time.time() - t0:.1f}s "
        f"(ratio {distinct * size / lens.sum():.2f}, {os.cpu_count()} host cores)")
    _save(path, blob=blob, offs=offs, lens=lens, crc=crc)
    return blob, offs, lens, crc


def build_corpus3(distinct: int, block: int):
    """`distinct` raw-LZMA2 streams of one block each (text_block(seed=i), preset 6), terminator stripped: the
    building blocks of config 3's single stream with a dictionary reset per block."""
    path = os.path.join(_cache_dir(), f"lzma_b200_corpus_lzma2_{distinct}x{block}_lc3lp0pb2_d8M_p6.npz")
    if os.path.exists(path):
        z = np.load(path)
        return z["blob"], z["offs"], z["lens"], z["crc"]
    from concurrent.futures import ProcessPoolExecutor
    from lzma_b200 import corpus as K
    t0 = time.time()
    with ProcessPoolExecutor(max_workers=min(os.cpu_count() or 1, distinct)) as ex:
        out = list(ex.map(_job_lzma2_block, [(i, block) for i in range(distinct)], chunksize=max(1, distinct // 64)))
    blob, offs, lens = _pack([o[0] for o in out])
    crc = np.array([o[1] for o in out], dtype=np.uint32)
    log(f"[bench] config 3 corpus: {distinct} LZMA2 blocks x {block} B in {time.time() - t0:.1f}s")
    _save(path, blob=blob, offs=offs, lens=lens, crc=crc)
    return blob, offs, lens, crc


def _job_lzma2_block(args):
    from lzma_b200 import corpus as K
    seed, block = args
    d = K.text_block(seed, block)
    s = K.compress_raw_lzma2(d)
    assert s[-1] == 0
    return s[:-1], zlib.crc32(d)


def build_corpus5(distinct: int, size: int):
    """`distinct` .lzma streams of varied compressibility (corpus.varied_block), liblzma preset 1."""
    path = os.path.join(_cache_dir(), f"lzma_b200_corpus_varied_{distinct}x{size}_lc3lp0pb2_d8M_p1.npz")
    if os.path.exists(path):
        z = np.load(path)
        return z["blob"], z["offs"], z["lens"], z["crc"]
    from lzma_b200 import corpus as K
    t0 = time.time()
    streams, crcs = K.build_varied_streams(distinct, size, seed0=0, preset=1)
    blob, offs, lens = _pack(streams)
    crc = np.array(crcs, dtype=np.uint32)
    log(f"[bench] config 5 corpus: {distinct} streams x {size} B in {time.time() - t0:.1f}s "
        f"(compressed/plain {lens.min() / size:.2f} .. {lens.max() / size:.2f})")
    _save(path, blob=blob, offs=offs, lens=lens, crc=crc)
    return blob, offs, lens, crc


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows: list[list[str]] = []
        self.proc = None
        self.th = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        def rd():
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        self.th = threading.Thread(target=rd, daemon=True)
        self.th.start()

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for k, nm in enumerate(names):
                if len(r) > 3 + k and r[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU leg
def cpu_pass(blob, offs, lens, sample_idx, size, threads):
    """One pass of the CPU oracle over the sampled streams; returns (seconds, n_bad)."""
    from oracle import oracle as O
    n = len(sample_idx)
    in_off = offs[sample_idx].astype(np.uint64)
    in_len = lens[sample_idx].astype(np.uint64)
    out = np.empty(n * size, dtype=np.uint8)
    out_off = (np.arange(n, dtype=np.uint64) * np.uint64(size))
    out_cap = np.full(n, size, dtype=np.uint64)
    t0 = time.perf_counter()
    bad, _ = O.lzma_alone_batch(blob, in_off, in_len, out, out_off, out_cap, threads)
    return time.perf_counter() - t0, bad


def cpu_sample(distinct: int, cores: int, streams: int = 1024):
    """Bounded sample of the workload for the CPU legs: 1-2 s of wall time on all cores (about 17 ms of one
    core per 1 MiB stream: the whole 1 024-stream workload is ~17 core-seconds; fewer streams on small hosts)."""
    n = max(32, min(streams, cores * 64))
    return np.arange(n) % distinct


def headline_config(args, world, distinct, lens):
    """The `config` object of the JSON line: identical for both arms (the driver compares them key by key)."""
    g_n = world * args.streams
    g_stream = np.arange(g_n) % distinct
    comp = int(lens[g_stream].sum())
    return {"workload": (f"{args.streams} independent .lzma streams x {args.size} B synthetic text-like "
                         f"(Zipf words), lc3 lp0 pb2, 8 MiB dict, liblzma preset 6, per GPU"),
            "units_total": g_n, "distinct_streams": distinct,
            "compressed_bytes_total": comp, "decompressed_bytes_total": g_n * args.size,
            "sharding": "LPT by compressed size over ranks, no collective",
            "l2": "per-step working set (compressed in + decoded out) exceeds the 126 MB L2"}


# ----------------------------------------------------------------------------- sharded device-resident job
class Env:
    """What the sub-benchmarks need from main(): torch, the process group, the library context."""
    def __init__(self, torch, dist, L, ctx, rank, world, local_rank):
        self.torch, self.dist, self.L, self.ctx = torch, dist, L, ctx
        self.rank, self.world, self.local_rank = rank, world, local_rank
        self.stream = torch.cuda.current_stream().cuda_stream or 1   # 0 would mean "the context's own stream"

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()

    def reduce(self, vals, op):
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=getattr(self.dist.ReduceOp, op))
        return t.tolist()


def run_sharded(env: Env, g_units, blob: np.ndarray, want_crc, steps: int, warmup: int, e2e_steps: int = 0,
                e2e_prepare=None):
    """A global batch `g_units` (ctypes array; in_off into `blob`, out_cap set, out_off ignored) sharded over the
    ranks by compressed size; this rank decodes its shard with everything resident in HBM.  Device-timed with CUDA
    events, max over ranks; every unit's decoded bytes verified against `want_crc[i]` by the on-device CRC-32.
    e2e_steps > 0: the shard once more through lzgpu_decode_batch with pinned host buffers (e2e_prepare() is called
    inside each timed call first: the host-side work a user of the public API does, e.g. scanning an LZMA2 stream)."""
    L, torch = env.L, env.torch
    from lzma_b200.batch import Unit
    n_all = len(g_units)
    shard = (C.c_int32 * n_all)()
    L.check(L.lib().lzgpu_shard_units(g_units, n_all, env.world, shard))
    mine = [i for i in range(n_all) if shard[i] == env.rank]
    n = len(mine)
    units = (Unit * max(n, 1))()
    off = 0
    for k, i in enumerate(mine):
        C.memmove(C.byref(units[k]), C.byref(g_units[i]), C.sizeof(Unit))
        units[k].out_off = off
        off = (off + int(units[k].out_cap) + 15) & ~15
    out_size = off + 16
    out_bytes = sum(int(units[k].out_cap) for k in range(n))
    comp_bytes = sum(int(units[k].in_len) for k in range(n))
    d_in = torch.from_numpy(blob).cuda()
    d_out = torch.empty(out_size, dtype=torch.uint8, device="cuda")
    plan = env.ctx.plan(units, blob.nbytes, out_size) if n else None

    def step():
        if plan:
            plan.launch(d_in.data_ptr(), d_out.data_ptr(), env.stream)

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    env.barrier()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    env.barrier()
    ms = e0.elapsed_time(e1) / steps
    verified = 0
    if plan:
        res, _ = plan.results()
        bad = [(mine[k], res[k].status, res[k].err_site) for k in range(n) if res[k].status != L.OK or res[k].bytes_out != units[k].out_cap]
        assert not bad, f"decode failed on units {bad[:5]}"
        crc = plan.crc32(d_out.data_ptr())
        for k, i in enumerate(mine):
            assert int(crc[k]) == int(want_crc[i]), f"unit {i} decoded wrongly (device CRC-32)"
        verified = n
    launches = plan.launch_count if plan else 0
    e2e_s = 0.0
    if e2e_steps:
        from lzma_b200.batch import pinned_empty
        h_in = pinned_empty(blob.nbytes)
        h_in[:blob.nbytes] = blob
        h_out = pinned_empty(out_size)
        for it in range(1 + e2e_steps):
            if it == 1:
                env.barrier()
                t0 = time.perf_counter()
            if e2e_prepare:
                e2e_prepare()
            if n:
                r2, _ = env.ctx.decode_batch(units, h_in[:blob.nbytes], h_out[:out_size])
        e2e_s = (time.perf_counter() - t0) / e2e_steps
        if n:
            assert all(r2[k].status == L.OK for k in range(n))
            for k in range(0, n, max(1, n // 128)):      # a sample on the host; the device CRC covered every unit above
                o = int(units[k].out_off)
                assert zlib.crc32(h_out[o:o + int(units[k].out_cap)]) == int(want_crc[mine[k]]), f"e2e: unit {mine[k]} decoded wrongly"
        del h_in, h_out
    if plan:
        plan.close()
    del d_in, d_out
    torch.cuda.empty_cache()
    ms_max, e2e_max = env.reduce([ms, e2e_s], "MAX")
    g_out, g_comp, g_ver, g_n = env.reduce([out_bytes, comp_bytes, verified, n], "SUM")
    n_min, = env.reduce([n], "MIN")
    n_max, = env.reduce([n], "MAX")
    return {"ms": ms_max, "value": g_out / (ms_max * 1e-3) / 1e9, "unit": "GB/s", "steps": steps, "warmup": warmup,
            "units_total": int(g_n), "units_per_gpu": [int(n_min), int(n_max)],
            "compressed_bytes_total": int(g_comp), "decompressed_bytes_total": int(g_out),
            "verified": f"{int(g_ver)} of {int(g_n)} units: status OK, declared size, on-device CRC-32 (lzgpu_plan_crc32) == plaintext CRC-32",
            "gpu_launches": steps * launches,
            "e2e": ({"value": g_out / e2e_max / 1e9, "unit": "GB/s", "ms": e2e_max * 1e3, "steps": e2e_steps,
                     "h2d_bytes_per_step": int(g_comp), "d2h_bytes_per_step": int(g_out)} if e2e_steps else None)}


def run_config3(env: Env, distinct: int, blocks: int, block: int, steps: int, warmup: int):
    """BASELINE config 3: ONE raw LZMA2 stream, dictionary reset every `block` bytes, decoded in parallel."""
    L = env.L
    from lzma_b200.batch import Unit
    if env.rank == 0:
        build_corpus3(distinct, block)
    env.barrier()
    blob, offs, lens, crc = build_corpus3(distinct, block)
    pick = np.arange(blocks) % distinct
    stream = np.concatenate([blob[offs[i]:offs[i] + lens[i]] for i in pick] + [np.zeros(1, dtype=np.uint8)])
    lib = L.lib()

    def scan():
        arr = (Unit * blocks)()
        total, sst = C.c_uint64(), C.c_int32()
        n = lib.lzgpu_scan_lzma2(stream.ctypes.data, stream.nbytes, 8 << 20, arr, blocks, C.byref(total), C.byref(sst))
        assert n == blocks and sst.value == L.OK and total.value == blocks * block, (n, sst.value, total.value)
        return arr

    t0 = time.perf_counter()
    g_units = scan()
    scan_ms = (time.perf_counter() - t0) * 1e3
    want = [int(crc[i]) for i in pick]
    r = run_sharded(env, g_units, stream, want, steps, warmup, e2e_steps=max(1, min(steps, 3)), e2e_prepare=scan)
    r["workload"] = (f"ONE raw LZMA2 stream, {blocks} x {block} B text-like blocks ({blocks * block / 2**30:.2f} GiB), dictionary "
                     f"reset per block, lc3 lp0 pb2, liblzma preset 6 ({distinct} distinct blocks); host scan "
                     f"(lzgpu_scan_lzma2) -> {blocks} units, sharded over the ranks")
    r["scaling"] = "strong"
    r["host_scan_ms"] = scan_ms
    r["e2e"]["api"] = "lzgpu_scan_lzma2 + lzgpu_decode_batch (pinned host buffers), scan inside the timed region"
    return r


def run_config5(env: Env, distinct: int, n_units: int, size: int, steps: int, warmup: int):
    """BASELINE config 5: fixed global batch, LPT-sharded over the ranks (strong scaling)."""
    L = env.L
    from lzma_b200.batch import Unit, parse_alone_header
    if env.rank == 0:
        build_corpus5(distinct, size)
    env.barrier()
    blob, offs, lens, crc = build_corpus5(distinct, size)
    # 16-byte aligned slots for the distinct streams, resident once; unit j decodes stream j % distinct into its own range
    slot = np.zeros(distinct, dtype=np.int64)
    pos = 0
    for i in range(distinct):
        slot[i] = pos
        pos = (pos + int(lens[i]) + 15) & ~15
    packed = np.zeros(pos + 16, dtype=np.uint8)
    proto = []
    for i in range(distinct):
        packed[slot[i]:slot[i] + lens[i]] = blob[offs[i]:offs[i] + lens[i]]
        st, u = parse_alone_header(blob[offs[i]:offs[i] + 13].tobytes())
        assert st == L.OK
        u.kind = L.KIND_LZMA1_ALONE
        u.in_off, u.in_len, u.out_cap = int(slot[i]), int(lens[i]), size
        proto.append(u)
    g_units = (Unit * n_units)()
    for j in range(n_units):
        C.memmove(C.byref(g_units[j]), C.byref(proto[j % distinct]), C.sizeof(Unit))
    want = [int(crc[j % distinct]) for j in range(n_units)]
    r = run_sharded(env, g_units, packed, want, steps, warmup)
    r["workload"] = (f"FIXED global batch of {n_units} .lzma streams x {size} B ({n_units * size / 2**30:.0f} GiB decoded), "
                     f"{distinct} distinct seeds of varied compressibility (text slices with 0-21% noise and 0-12% runs/records, "
                     f"compressed/plain {lens.min() / size:.2f}..{lens.max() / size:.2f}), lc3 lp0 pb2, 8 MiB dict, liblzma preset 1; "
                     f"each distinct stream resident once and decoded into {n_units // distinct} separate output ranges; "
                     f"LPT-sharded by compressed size over the ranks, no collective")
    r["scaling"] = "strong"
    r["distinct_streams"] = distinct
    r.pop("e2e")
    r["e2e"] = None
    r["e2e_note"] = "device-resident only: 64 GiB of output is verified where it lies (on-device CRC-32), not copied back"
    return r


def run_config4(env: Env):
    """BASELINE config 4: the mixed batch, every unit against the oracle (rank 0 only: a parity case with a clock)."""
    L, torch = env.L, env.torch
    from lzma_b200 import batch as B
    from lzma_b200 import corpus as K
    from oracle import oracle as O                  # the checker, never the thing measured
    items = K.mixed_batch(assets_dir=os.path.join(ROOT, "tests", "golden", "ref_assets"))
    units, owner, blobs = [], [], []
    in_off = out_off = 0

    for idx, it in enumerate(items):
        d = it["data"]
        if it["kind"] == "lzma2":
            us, total, _sst = B.scan_lzma2(d, it["dict"])
            for u in us:
                u.in_off += in_off
                u.out_off += out_off
                units.append(u)
                owner.append(idx)
            cap = max(int(total), 16)
        else:
            u = L.Unit()
            if it["kind"] == "alone":
                _st, u = B.parse_alone_header(d)
                u.kind = L.KIND_LZMA1_ALONE
            else:
                u.kind = L.KIND_LZMA1_RAW
                u.lc, u.lp, u.pb, u.dict_size, u.unpack_size = it["lc"], it["lp"], it["pb"], it["dict"], it["unpack"]
            u.in_off, u.in_len, u.out_off, u.out_cap = in_off, len(d), out_off, it["cap"]
            units.append(u)
            owner.append(idx)
            cap = it["cap"]
        blobs.append((in_off, d))
        in_off = (in_off + len(d) + 15) & ~15
        out_off = (out_off + cap + 15) & ~15
    in_size, out_size = in_off + 16, out_off + 16
    h_in = B.pinned_empty(in_size)
    h_in[:] = 0
    for o, d in blobs:
        h_in[o:o + len(d)] = np.frombuffer(d, dtype=np.uint8)
    h_out = B.pinned_empty(out_size)
    n = len(units)
    arr = (L.Unit * n)(*units)
    res = None
    t_best = None
    for it in range(4):                              # 1 warm-up + 3 timed calls, host buffers (the public API)
        t0 = time.perf_counter()
        res, st = env.ctx.decode_batch(arr, h_in[:in_size], h_out[:out_size])
        dt = time.perf_counter() - t0
        if it:
            t_best = dt if t_best is None else min(t_best, dt)
    # oracle: per .lzma / headerless unit; per LZMA2 stream (its units' outputs are consecutive)
    mismatches, kinds, dec_bytes = [], {}, 0
    k = 0
    for idx, it in enumerate(items):
        mine = [j for j in range(k, n) if owner[j] == idx] if it["kind"] == "lzma2" else [k]
        k = mine[-1] + 1
        d = it["data"]
        if it["kind"] == "alone":
            want = O.lzma_alone(d, it["cap"])
        elif it["kind"] == "raw":
            want = O.lzma_raw(d, it["lc"], it["lp"], it["pb"], it["dict"], it["unpack"], it["cap"])
        else:
            want = O.lzma2(d, it["dict"], it["cap"] + (1 << 20))
        if it["kind"] == "lzma2":
            got_status, got_site, got = L.OK, 0, b""
            for j in mine:
                got += h_out[units[j].out_off:units[j].out_off + res[j].bytes_out].tobytes()
                if res[j].status != L.OK:
                    got_status, got_site = res[j].status, res[j].err_site
                    break
        else:
            r = res[mine[0]]
            got_status, got_site = r.status, r.err_site
            got = h_out[units[mine[0]].out_off:units[mine[0]].out_off + r.bytes_out].tobytes()
        ok = got_status == want.status
        if ok and want.status in (O.OK, O.OK_INPUT_EXHAUSTED):
            ok = got == want.data
        elif ok and want.status == O.RESULT_ERROR and it["kind"] != "lzma2":
            ok = got_site == want.err_site
        if not ok:
            mismatches.append((it["name"], got_status, got_site, want.status, want.err_site))
        kinds[want.status_name] = kinds.get(want.status_name, 0) + 1
        dec_bytes += len(got)
    assert not mismatches, f"config 4: GPU != oracle on {mismatches[:5]}"
    return {"workload": (f"mixed batch of {len(items)} streams = {n} units in ONE lzgpu_decode_batch call: .lzma with all 75 "
                         "lc/lp/pb of liblzma + 8 re-labelled beyond lc+lp=4 (literal tables in HBM), EOS / size / both, headerless "
                         "LZMA1, raw LZMA2 with dictionary resets and uncompressed chunks, incompressible data, the reference's "
                         "test assets, 60+ corrupt / truncated / bad-header streams"),
            "units_total": n, "streams": len(items), "ms": t_best * 1e3, "value": dec_bytes / t_best / 1e9, "unit": "GB/s",
            "timing": "end to end through lzgpu_decode_batch with pinned host buffers (best of 3 calls); a parity case: the batch is as slow as its longest unit -- the reference's own asset randomfile.dat.lzma, 1 MiB of incompressible data in ONE stream (250 ms on one warp); the table classes' launches run side by side",
            "decompressed_bytes_total": dec_bytes, "kernel_ms": st.kernel_ms, "gpu_launches": int(st.launches),
            "outcomes_by_oracle_status": kinds,
            "verified": f"status, error site (LZMA1) and decoded bytes of all {len(items)} streams == oracle; 0 mismatches"}


# ----------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--streams", type=int, default=1024, help="units per GPU (headline workload)")
    ap.add_argument("--size", type=int, default=STREAM_SIZE)
    ap.add_argument("--distinct", type=int, default=0, help="distinct streams to generate (0: auto)")
    ap.add_argument("--configs", default="3,4,5", help="BASELINE configurations measured as sub-records beside the headline (config 2); '' for none")
    ap.add_argument("--c5-units", type=int, default=16384)
    ap.add_argument("--c5-size", type=int, default=4 << 20)
    ap.add_argument("--c5-distinct", type=int, default=1024)
    ap.add_argument("--c5-steps", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    extra = [c for c in args.configs.split(",") if c]

    if args.gpus > 1 and "WORLD_SIZE" not in os.environ and args.impl == "ours":
        # launched without torchrun: re-exec as one process per GPU, the contract's launch form
        import socket
        s = socket.socket()
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
        s.close()
        os.execvp(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                                   f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
                                   "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:])
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cores = os.cpu_count() or 1
    distinct = args.distinct or (args.streams if cores >= 32 else min(args.streams, 256))

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        blob, offs, lens, _ = build_corpus(distinct, args.size)
        idx = cpu_sample(distinct, cores, args.streams)
        for _ in range(args.warmup):
            cpu_pass(blob, offs, lens, idx, args.size, cores)
        t = 0.0
        for _ in range(args.steps):
            dt, bad = cpu_pass(blob, offs, lens, idx, args.size, cores)
            assert bad == 0, "CPU oracle failed on the corpus"
            t += dt
        gbs = len(idx) * args.size * args.steps / t / 1e9
        sample = f"{len(idx)} of the workload's streams per step, one stream per task over {cores} threads"
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": gbs, "unit": "GB/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": headline_config(args, max(args.gpus, 1), distinct, lens),
            "note": "CPU arm: C restatement of the reference's Go decoder (oracle/); Go toolchain absent, reference itself not runnable",
            "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }))
        return 0

    # ------------------------------------------------------------------ our arm
    if not os.path.exists(os.path.join(ROOT, "lzma_b200", "liblzgpu.so")):
        import __graft_entry__
        __graft_entry__.build()      # compile the CUDA extension in-tree (nvcc cross-compiles anywhere)
    import torch
    import torch.distributed as dist
    from lzma_b200 import _lib as L
    from lzma_b200.batch import Context, Unit, parse_alone_header

    if L.lib().lzgpu_device_count() <= 0:
        raise SystemExit("bench.py: no CUDA device; lzma_b200 has no CPU decode path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        # NCCL prints its version banner to STDOUT (at NCCL_DEBUG=VERSION / WARN / INFO); stdout carries the one
        # JSON line, so the communicator is created (and used once) with fd 1 pointing at stderr
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            import datetime
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=datetime.timedelta(minutes=30))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    def barrier():
        if world > 1:
            dist.barrier()

    # corpus: rank 0 builds (or finds) the cache, everybody loads it
    if rank == 0:
        build_corpus(distinct, args.size)
    barrier()
    blob, offs, lens, crc = build_corpus(distinct, args.size)

    # global batch of world x streams units, sharded over ranks by compressed size
    g_n = world * args.streams
    g_stream = np.arange(g_n) % distinct
    g_units = (Unit * g_n)()
    for j in range(g_n):
        g_units[j].in_len = int(lens[g_stream[j]])
    shard = (C.c_int32 * g_n)()
    L.check(L.lib().lzgpu_shard_units(g_units, g_n, world, shard))
    mine = [j for j in range(g_n) if shard[j] == rank]
    n = len(mine)

    # lay the rank's inputs into one buffer (16-byte aligned slots) and size the output
    units = (Unit * n)()
    in_off = 0
    in_offs = []
    for k, j in enumerate(mine):
        s = int(g_stream[j])
        st, u = parse_alone_header(blob[offs[s]:offs[s] + 13].tobytes())
        assert st == L.OK
        u.kind = L.KIND_LZMA1_ALONE
        u.in_off, u.in_len = in_off, int(lens[s])
        u.out_off, u.out_cap = k * args.size, args.size
        units[k] = u
        in_offs.append(in_off)
        in_off = (in_off + int(lens[s]) + 15) & ~15
    in_size, out_size = in_off + 16, n * args.size + 16
    h_in = torch.empty(in_size, dtype=torch.uint8, pin_memory=True)
    h_out = torch.empty(out_size, dtype=torch.uint8, pin_memory=True)
    hin = h_in.numpy()
    for k, j in enumerate(mine):
        s = int(g_stream[j])
        hin[in_offs[k]:in_offs[k] + lens[s]] = blob[offs[s]:offs[s] + lens[s]]
    comp_bytes = int(sum(int(lens[g_stream[j]]) for j in mine))
    out_bytes = n * args.size

    ctx = Context([local_rank])
    env = Env(torch, dist, L, ctx, rank, world, local_rank)
    d_in = h_in.cuda(non_blocking=False)
    d_out = torch.empty(out_size, dtype=torch.uint8, device="cuda")
    plan = ctx.plan(units, in_size, out_size)
    stream = env.stream

    def step():
        plan.launch(d_in.data_ptr(), d_out.data_ptr(), stream)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    res, _ = plan.results()
    bad = [(k, res[k].status) for k in range(n) if res[k].status != L.OK or res[k].bytes_out != args.size]
    assert not bad, f"decode failed on units {bad[:5]}"

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    torch.cuda.synchronize()
    ev[0].record()
    for i in range(args.steps):
        step()
        ev[i + 1].record()
    torch.cuda.synchronize()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    total_ms = ev[0].elapsed_time(ev[-1])
    step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    res, kstats = plan.results()
    bad = [(k, res[k].status) for k in range(n) if res[k].status != L.OK or res[k].bytes_out != args.size]
    assert not bad, f"decode failed on units {bad[:5]}"

    # bit-exact check of the timed output: CRC32 of every decoded unit whose plaintext CRC is cached, computed
    # where the bytes lie (lzgpu_plan_crc32) and again on the host
    dev_crc = plan.crc32(d_out.data_ptr())
    for k, j in enumerate(mine):
        s = int(g_stream[j])
        assert int(dev_crc[k]) == int(crc[s]), f"unit {k} (stream {s}) decoded wrongly (device CRC)"
    out_host = d_out.cpu().numpy()
    checked = 0
    for k, j in enumerate(mine):
        s = int(g_stream[j])
        got = zlib.crc32(out_host[k * args.size:(k + 1) * args.size])
        assert got == int(crc[s]), f"unit {k} (stream {s}) decoded wrongly"
        checked += 1
    del out_host
    launch_count = plan.launch_count

    # ---- e2e: host buffers through the public batch call, --steps calls ----
    e2e = None
    if not args.no_e2e:
        hout = h_out.numpy()
        for _ in range(2):
            ctx.decode_batch(units, hin, hout)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            r2, st2 = ctx.decode_batch(units, hin, hout)
        t_e2e = (time.perf_counter() - t0) / args.steps
        assert all(r2[k].status == L.OK for k in range(n))
        for k, j in enumerate(mine):     # every unit of the end-to-end output, too
            s = int(g_stream[j])
            assert zlib.crc32(hout[units[k].out_off:units[k].out_off + args.size]) == int(crc[s]), f"e2e: unit {k} decoded wrongly"
        e2e = {"t": t_e2e, "kernel_ms": st2.kernel_ms, "h2d_ms": st2.h2d_ms, "d2h_ms": st2.d2h_ms, "launches": int(st2.launches)}
    plan.close()
    del d_in, d_out, h_in, h_out
    torch.cuda.empty_cache()

    # ---- max over ranks ----
    total_ms_max, e2e_t_max = env.reduce([total_ms, e2e["t"] if e2e else 0.0], "MAX")
    g_out, g_comp = env.reduce([float(out_bytes), float(comp_bytes)], "SUM")

    # ---- the other BASELINE configurations, same run ----
    sub = {}
    if "3" in extra:
        sub["config3"] = run_config3(env, distinct, 1024, 1 << 20, steps=max(1, min(args.steps, 3)), warmup=2)
    if "4" in extra:
        if rank == 0:
            sub["config4"] = run_config4(env)
        barrier()
    if "5" in extra:
        sub["config5"] = run_config5(env, args.c5_distinct, args.c5_units, args.c5_size, steps=args.c5_steps, warmup=1)

    if rank == 0:
        ms_per_step = total_ms_max / args.steps
        value = g_out / (ms_per_step * 1e-3) / 1e9
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "B200_PROFILING.md fallback 6650 (of fallback)"
        kernel_ms = statistics.mean(step_ms)  # rank 0's launches, one kernel per step
        achieved = (comp_bytes + out_bytes) / (kernel_ms * 1e-3) / 1e9
        traffic, traffic_note = None, "no ncu --set full capture of this kernel version in profiles/traffic.json"
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            if tj.get("kernel_tag") == KERNEL_TAG:
                traffic, traffic_note = tj.get("dram_bytes_per_launch"), tj.get("source")
        except Exception:
            pass
        config = headline_config(args, world, distinct, lens)
        assert config["compressed_bytes_total"] == int(g_comp) and config["decompressed_bytes_total"] == int(g_out)
        line = {
            "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": config,
            "verified": f"CRC32 of {checked} decoded units per rank vs plaintext, on the device (lzgpu_plan_crc32) and on the host, for the device-resident and the end-to-end output; status OK + size for all",
            "clocks": clocks,
            "gpu_launches": args.steps * launch_count,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_note, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": comp_bytes + out_bytes,
                         "kernel": KERNEL_NAME, "kernel_ms": kernel_ms,
                         "note": "latency-bound: one serial range-decoder chain per unit; see DESIGN.md"},
        }
        if e2e:
            line["e2e"] = {"value": g_out / e2e_t_max / 1e9, "unit": "GB/s", "h2d_bytes_per_step": int(g_comp),
                           "d2h_bytes_per_step": int(g_out), "ms_per_step": e2e_t_max * 1e3, "steps": args.steps,
                           "gpu_launches_per_step": e2e["launches"],
                           "rank0_breakdown_ms": {"h2d": e2e["h2d_ms"], "kernel": e2e["kernel_ms"], "d2h": e2e["d2h_ms"]},
                           "api": "lzgpu_decode_batch (pinned host buffers; the units read the compressed input from host memory over PCIe while they decode and write every finished 64 KiB block of output into the caller's buffer themselves; a device falls back to copy-engine transfers overlapped with the kernel when the host cannot keep up)"}
        line.update(sub)
        if not args.no_cpu_baseline:
            idx = cpu_sample(distinct, cores, args.streams)
            cpu_pass(blob, offs, lens, idx[:max(1, len(idx) // 4)], args.size, cores)  # warm
            dt, nbad = cpu_pass(blob, offs, lens, idx, args.size, cores)
            assert nbad == 0
            line["cpu_baseline"] = {"value": len(idx) * args.size / dt / 1e9, "unit": "GB/s", "cores": cores, "kind": "port",
                                    "sample": f"{len(idx)} of the workload's streams, one stream per task over {cores} threads, {dt:.2f}s"}
        print(json.dumps(line))
    barrier()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
