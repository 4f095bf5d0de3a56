#!/bin/bash
# candidate build (LZGPU_LIB) : whole GPU tier, then timing against the in-tree library
mkdir -p gpurun_out
LZGPU_LIB=$PWD/lzma_b200/ab/lib_$1.so timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for n in intree $1; do
  echo "== $n"
  if [ $n = intree ]; then unset LZGPU_LIB; else export LZGPU_LIB=$PWD/lzma_b200/ab/lib_$n.so; fi
  timeout 600 python scripts/bench_corpora.py --shapes ${SHAPES:-text:148,text:1024,text:2072,mixed:1024} 2>&1 | grep -v Warning
  timeout 600 python scripts/bench_corpora.py --lzma2 2>&1 | grep -v Warning
done
