#!/bin/bash
# candidate (lib_nowrap) against the current build (lib_sched2): GPU tier on the candidate, then timing of both
mkdir -p gpurun_out
LZGPU_LIB=$PWD/lzma_b200/ab/lib_nowrap.so timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for rep in 1 2; do
for n in sched2 nowrap; do
  echo "== $n (rep $rep)"
  LZGPU_LIB=$PWD/lzma_b200/ab/lib_$n.so timeout 600 python scripts/bench_corpora.py --shapes text:148,text:1024,text:2072,random:1024,mixed:1024 2>&1 | grep -v Warning
done
done
