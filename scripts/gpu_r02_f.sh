#!/bin/bash
mkdir -p gpurun_out
free -g | head -2
python - <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import bench
blob, offs, lens, _ = bench.build_corpus3(256, 1 << 20)
with open("/tmp/s16.lzma2", "wb") as f:
    for i in range(16384):
        k = i % 256
        f.write(blob[offs[k]:offs[k] + lens[k]].tobytes())
    f.write(b"\0")
PY
tests/cpp/_build/reader2_bench /tmp/s16.lzma2 1048576 1073741824 2>&1 | tail -1 | tee gpurun_out/r02_reader2_16gib.json
rm -f /tmp/s16.lzma2
which compute-sanitizer
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "alone_cases or lzma2_uncompressed or long_and_overlapping or lzma2_cases" > gpurun_out/r02_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -12 gpurun_out/r02_sanitizer_memcheck.log
