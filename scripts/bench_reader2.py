#!/usr/bin/env python
"""Streaming facade throughput: builds ONE raw LZMA2 stream of `--gib` GiB (1 MiB text-like blocks,
dictionary reset per block: bench.py's config-3 corpus cycled), writes it to local disk and runs
tests/cpp/_build/reader2_bench on it (NewReader2 + Read, with and without decode-ahead); then the same stream through
the Python mirror (lzma_b200.reader2) with large reads.  Prints JSON lines."""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--gib", type=float, default=4.0)
ap.add_argument("--distinct", type=int, default=256)
args = ap.parse_args()
blob, offs, lens, _ = bench.build_corpus3(args.distinct, 1 << 20)
blocks = int(args.gib * 1024)
path = "/tmp/lzma_b200_reader2_stream.lzma2"
with open(path, "wb") as f:
    for i in range(blocks):
        k = i % args.distinct
        f.write(blob[offs[k]:offs[k] + lens[k]].tobytes())
    f.write(b"\0")
exe = os.path.join(ROOT, "tests", "cpp", "_build", "reader2_bench")
for bufsize, wave in ((32 << 10, 1 << 30), (1 << 20, 1 << 30), (1 << 20, 2 << 30), (1 << 20, 256 << 20)):
    print(subprocess.run([exe, path, str(bufsize), str(wave)], capture_output=True, text=True).stdout.strip(), flush=True)

import io  # noqa: E402
from lzma_b200.reader2 import NewReader2  # noqa: E402
data = open(path, "rb").read()
buf = np.empty(64 << 20, dtype=np.uint8)
for ahead in (True, True, False):
    r, err = NewReader2(io.BytesIO(data), 8 << 20)
    assert err is None
    r.decode_ahead = ahead
    t0 = time.perf_counter()
    total = 0
    while True:
        n, err = r.Read(buf)
        total += n
        if err is not None:
            break
    dt = time.perf_counter() - t0
    print(json.dumps({"what": "lzma_b200.reader2.NewReader2 + Read (Python mirror, 64 MiB reads)", "decode_ahead": ahead,
                      "decoded_bytes": total, "s": round(dt, 4), "GBps": round(total / dt / 1e9, 3)}), flush=True)
os.unlink(path)
