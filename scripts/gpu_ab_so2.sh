#!/bin/bash
# A/B of two builds (lzma_b200/liblzgpu_head.so vs liblzgpu.so): parity tests on the candidate, then the lone-warp
# shape and the bench shape (text), and the incompressible lone-warp shape, for both.
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log; else timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "fuzz or alone_cases or lzma2_cases or reference_assets" 2>&1 | tail -2; fi
cp lzma_b200/liblzgpu.so /tmp/cand.so
for rep in ${REPS:-1 2}; do
  for w in head cand; do
    if [ $w = head ]; then cp lzma_b200/liblzgpu_head.so lzma_b200/liblzgpu.so; else cp /tmp/cand.so lzma_b200/liblzgpu.so; fi
    echo "== $w"; timeout 600 python scripts/bench_corpora.py --quick 2>&1 | tail -2
    if [ $rep = 1 ]; then timeout 600 python scripts/bench_corpora.py --quick-random 2>&1 | tail -2; fi
  done
done
cp /tmp/cand.so lzma_b200/liblzgpu.so
