#!/bin/bash
for rep in 1 2; do
for n in "$@"; do
  echo "== $n"
  LZGPU_LIB=$PWD/lzma_b200/ab/lib_$n.so timeout 600 python scripts/bench_corpora.py --lzma2 2>&1 | grep -v Warning | head -1
done
done
