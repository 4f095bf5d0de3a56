#!/bin/bash
mkdir -p gpurun_out
LZGPU_LIB=$PWD/lzma_b200/ab/lib_pf1.so timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
for rep in 1 2; do
for n in pf0 pf1; do
  echo "== $n (rep $rep)"
  LZGPU_LIB=$PWD/lzma_b200/ab/lib_$n.so timeout 600 python scripts/bench_corpora.py --shapes text:148,text:1024,text:2072,mixed:1024 2>&1 | grep -v Warning
done
done
for rot in 8 32; do
  echo "== rotate $rot"
  LZGPU_ROTATE=$rot LZGPU_LIB=$PWD/lzma_b200/ab/lib_pf0.so timeout 600 python scripts/bench_corpora.py --shapes text:2072,mixed:1024,mixed:2072 2>&1 | grep -v Warning
done
