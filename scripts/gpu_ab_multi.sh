#!/bin/bash
# A/B of several builds of the library (lzma_b200/ab/lib_*.so, selected through LZGPU_LIB): a parity subset on each,
# then the shapes in $SHAPES, device-timed.  usage: gpu_ab_multi.sh name1 name2 ...   (REPS=2 repeats the timing)
mkdir -p gpurun_out
SHAPES=${SHAPES:-text:148,text:1024,random:148,random:1024,mixed:1024}
for n in "$@"; do
  export LZGPU_LIB=$PWD/lzma_b200/ab/lib_$n.so
  echo "== $n parity"; timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "${PARITY:-fuzz or alone_cases or lzma2_cases or reference_assets or encoder_cases}" 2>&1 | tail -2
done
for rep in ${REPS:-1}; do
  for n in "$@"; do
    export LZGPU_LIB=$PWD/lzma_b200/ab/lib_$n.so
    echo "== $n timing (rep $rep)"; timeout 900 python scripts/bench_corpora.py --shapes $SHAPES 2>&1 | grep -v Warning | tee -a gpurun_out/ab_$n.jsonl
  done
done
