#!/bin/bash
# scheduler: exchange policy A/B (LZGPU_ROTATE_MODE 0 = towards "more work left on the less crowded sub-partition", 1 = blind)
mkdir -p gpurun_out
export LZGPU_LIB=$PWD/lzma_b200/ab/lib_sched2.so
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
for mode in ${MODES:-0 1}; do
  echo "== LZGPU_ROTATE_MODE=$mode"
  LZGPU_ROTATE_MODE=$mode timeout 600 python scripts/bench_corpora.py --shapes ${SHAPES:-text:1024,text:2072,mixed:1024,mixed:2072,random:1024} 2>&1 | grep -v Warning
  for n in ${C5:-2048}; do
  LZGPU_ROTATE_MODE=$mode timeout 600 python bench.py --configs 5 --c5-units $n --no-e2e --no-cpu-baseline --steps 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=d['config5']; print('config5 $n units:', round(c['ms'],1), 'ms', round(c['value'],3), 'GB/s', 'headline', d['ms_per_step'])"
  done
done
