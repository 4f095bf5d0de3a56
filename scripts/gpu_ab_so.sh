#!/bin/bash
# A/B of two builds of liblzgpu.so: lzma_b200/liblzgpu_head.so (baseline) vs lzma_b200/liblzgpu.so (candidate).
# Parity tests on the candidate, then lone-warp latency (148 units) and the bench shape (1024) for both, twice.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
cp lzma_b200/liblzgpu.so /tmp/cand.so
for rep in 1 2; do
  for w in head cand; do
    if [ $w = head ]; then cp lzma_b200/liblzgpu_head.so lzma_b200/liblzgpu.so; else cp /tmp/cand.so lzma_b200/liblzgpu.so; fi
    echo "== $w"; timeout 600 python scripts/bench_corpora.py --quick 2>&1 | tail -2
  done
done
cp /tmp/cand.so lzma_b200/liblzgpu.so
