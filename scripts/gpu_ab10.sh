#!/bin/bash
LZGPU_LIB=$PWD/lzma_b200/ab/lib_$1.so timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
for rep in 1 2; do
for n in "$@"; do
  echo "== $n"
  LZGPU_LIB=$PWD/lzma_b200/ab/lib_$n.so timeout 600 python scripts/bench_corpora.py --shapes text:148,text:1024,text:2072,random:1024 2>&1 | grep -v Warning
  LZGPU_LIB=$PWD/lzma_b200/ab/lib_$n.so timeout 600 python scripts/bench_corpora.py --lzma2 2>&1 | grep -v Warning | head -1
done
done
