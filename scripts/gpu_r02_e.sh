#!/bin/bash
mkdir -p gpurun_out
tests/cpp/_build/reader_test tests/golden/ref_assets | tail -4
python - <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import bench
blob, offs, lens, _ = bench.build_corpus3(256, 1 << 20)
with open("/tmp/s.lzma2", "wb") as f:
    for i in range(4096):
        k = i % 256
        f.write(blob[offs[k]:offs[k] + lens[k]].tobytes())
    f.write(b"\0")
PY
LZMA_READER_TRACE=1 tests/cpp/_build/reader2_bench /tmp/s.lzma2 1048576 1073741824 2>&1 | tail -14
tests/cpp/_build/reader2_bench /tmp/s.lzma2 32768 1073741824 2>&1 | tail -1
tests/cpp/_build/reader2_bench /tmp/s.lzma2 1048576 536870912 2>&1 | tail -1
