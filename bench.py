#!/usr/bin/env python
"""bench.py -- decompressed GB/s of the batch LZMA decode path on B200.

Workload (BASELINE.json configs[1]): a batch of 1024 independent .lzma streams x 1 MiB
of synthetic text-like data, lc3 lp0 pb2, 8 MiB dictionary, compressed by liblzma
preset 6, on one GPU.  With N GPUs (one process per GPU under torchrun) the global
batch is N x 1024 units, sharded over ranks by compressed size (lzgpu_shard_units);
there is no data-path collective (weak scaling).

A step = one pass of the decode path over the rank's units:
  value  : inputs and outputs resident in HBM, CUDA events around the launches
  e2e    : the same batch through lzgpu_decode_batch with pinned HOST buffers, H2D of the
           compressed input and D2H of the decoded output inside the timed region
  cpu_baseline / --impl reference : the C restatement of the reference's decoder
           (oracle/, kind "port": the Go toolchain is absent so the reference itself
           cannot run) on the box's host cores, bounded sample of the same streams.
"""
from __future__ import annotations

import argparse
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


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ----------------------------------------------------------------------------- corpus
def corpus_path(distinct: int, size: int) -> str:
    return os.path.join(os.environ.get("LZMA_B200_CACHE", "/tmp"), f"lzma_b200_corpus_text_{distinct}x{size}_lc3lp0pb2_d8M_p6.npz")


def build_corpus(distinct: int, size: int):
    """`distinct` .lzma streams (stream i: text_block(seed=i)), cached on local disk."""
    path = corpus_path(distinct, size)
    if os.path.exists(path):
        z = np.load(path)
        return z["blob"], z["offs"], z["lens"], z["crc"]
    from lzma_b200 import corpus as K
    t0 = time.time()
    streams, crcs = K.build_alone_streams(distinct, size, seed0=0, with_crc=True)
    crc = np.array(crcs, dtype=np.uint32)
    lens = np.array([len(s) for s in streams], dtype=np.int64)
    offs = np.zeros(distinct, dtype=np.int64)
    np.cumsum(lens[:-1], out=offs[1:])
    blob = np.frombuffer(b"".join(streams), dtype=np.uint8)
    log(f"[bench] built {distinct} streams x {size} B in {time.time() - t0:.1f}s "
        f"(ratio {distinct * size / lens.sum():.2f}, {os.cpu_count()} host cores)")
    try:
        tmp = path + f".{os.getpid()}.tmp.npz"
        np.savez(tmp, blob=blob, offs=offs, lens=lens, crc=crc)
        os.replace(tmp, path)
    except OSError:
        pass
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


# ----------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--streams", type=int, default=1024, help="units per GPU")
    ap.add_argument("--size", type=int, default=STREAM_SIZE)
    ap.add_argument("--distinct", type=int, default=0, help="distinct streams to generate (0: auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

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
    workload = (f"{args.streams} independent .lzma streams x {args.size} B synthetic text-like "
                f"(Zipf words), lc3 lp0 pb2, 8 MiB dict, liblzma preset 6, per GPU")

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
            "config": {"workload": workload, "note": "CPU arm: C restatement of the reference's Go decoder (oracle/); "
                       "Go toolchain absent, reference itself not runnable"},
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
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
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
    import ctypes as C
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
    d_in = h_in.cuda(non_blocking=False)
    d_out = torch.empty(out_size, dtype=torch.uint8, device="cuda")
    plan = ctx.plan(units, in_size, out_size)
    # torch's default stream has handle 0, which the C ABI reads as "use the context's own stream";
    # cudaStreamLegacy (0x1) names the same stream explicitly, so torch's events bracket the kernels.
    stream = torch.cuda.current_stream().cuda_stream or 1

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
        if s < len(crc):
            assert int(dev_crc[k]) == int(crc[s]), f"unit {k} (stream {s}) decoded wrongly (device CRC)"
    out_host = d_out.cpu().numpy()
    checked = 0
    for k, j in enumerate(mine):
        s = int(g_stream[j])
        if s < len(crc):
            got = zlib.crc32(out_host[k * args.size:(k + 1) * args.size])
            assert got == int(crc[s]), f"unit {k} (stream {s}) decoded wrongly"
            checked += 1
    del out_host

    # ---- e2e: host buffers through the public batch call ----
    e2e = None
    if not args.no_e2e:
        hout = h_out.numpy()
        for _ in range(2):
            ctx.decode_batch(units, hin, hout)
        e2e_steps = max(1, min(args.steps, 3))
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            r2, st2 = ctx.decode_batch(units, hin, hout)
        t_e2e = (time.perf_counter() - t0) / e2e_steps
        assert all(r2[k].status == L.OK for k in range(n))
        for k, j in enumerate(mine):     # every unit of the end-to-end output, too
            s = int(g_stream[j])
            if s < len(crc):
                assert zlib.crc32(hout[units[k].out_off:units[k].out_off + args.size]) == int(crc[s]), f"e2e: unit {k} decoded wrongly"
        e2e = {"t": t_e2e, "h2d": comp_bytes, "d2h": out_bytes, "kernel_ms": st2.kernel_ms, "h2d_ms": st2.h2d_ms, "d2h_ms": st2.d2h_ms}

    # ---- max over ranks ----
    t = torch.tensor([total_ms, e2e["t"] if e2e else 0.0], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(out_bytes), float(comp_bytes)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    total_ms_max, e2e_t_max = t.tolist()
    g_out, g_comp = tot.tolist()

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
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("dram_bytes_per_launch")
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": workload, "units_total": g_n, "distinct_streams": distinct,
                       "compressed_bytes_total": int(g_comp), "decompressed_bytes_total": int(g_out),
                       "sharding": "LPT by compressed size over ranks, no collective",
                       "l2": "per-step working set (compressed in + decoded out) exceeds the 126 MB L2",
                       "verified": f"CRC32 of {checked} decoded units vs plaintext, on the device (lzgpu_plan_crc32) and on the host, for the device-resident and the end-to-end output; status OK + size for all"},
            "clocks": clocks,
            "gpu_launches": args.steps * plan.launch_count,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": comp_bytes + out_bytes,
                         "kernel": "lzgpu_decode_kernel<false>", "kernel_ms": kernel_ms,
                         "note": "latency-bound: one serial range-decoder chain per unit; see DESIGN.md"},
        }
        if e2e:
            line["e2e"] = {"value": g_out / e2e_t_max / 1e9, "unit": "GB/s", "h2d_bytes_per_step": int(g_comp),
                           "d2h_bytes_per_step": int(g_out), "ms_per_step": e2e_t_max * 1e3,
                           "rank0_breakdown_ms": {"h2d": e2e["h2d_ms"], "kernel": e2e["kernel_ms"], "d2h": e2e["d2h_ms"]},
                           "api": "lzgpu_decode_batch (pinned host buffers; the units read the compressed input from host memory over PCIe while they decode, finished 64 KiB output blocks are copied out while the kernel runs)"}
        if world == 1 and not args.no_cpu_baseline:
            idx = cpu_sample(distinct, cores, args.streams)
            cpu_pass(blob, offs, lens, idx[:max(1, len(idx) // 4)], args.size, cores)  # warm
            dt, nbad = cpu_pass(blob, offs, lens, idx, args.size, cores)
            assert nbad == 0
            line["cpu_baseline"] = {"value": len(idx) * args.size / dt / 1e9, "unit": "GB/s", "cores": cores, "kind": "port",
                                    "sample": f"{len(idx)} of the workload's streams, one stream per task over {cores} threads, {dt:.2f}s"}
        print(json.dumps(line))
    plan.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
