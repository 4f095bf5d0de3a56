#!/bin/bash
# SM-resident scheduler, second pass: parity subset, then the shapes with a tail (heterogeneous units)
mkdir -p gpurun_out
export LZGPU_LIB=$PWD/lzma_b200/ab/lib_sched.so
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
for rot in ${ROTS:-16}; do
  echo "== LZGPU_ROTATE=$rot"
  LZGPU_ROTATE=$rot timeout 600 python scripts/bench_corpora.py --shapes ${SHAPES:-text:1024,text:2072,mixed:1024,mixed:2072} 2>&1 | grep -v Warning
  for n in ${C5:-2048}; do
  LZGPU_ROTATE=$rot timeout 600 python bench.py --configs 5 --c5-units $n --no-e2e --no-cpu-baseline --steps 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=d['config5']; print('config5 $n units:', round(c['ms'],1), 'ms', round(c['value'],3), 'GB/s', 'headline', d['ms_per_step'])"
  done
done
