#!/bin/bash
# A/B of builds on the bench line itself (device-timed and end to end) and two shapes; full GPU tier on the first
mkdir -p gpurun_out
LZGPU_LIB=$PWD/lzma_b200/ab/lib_$1.so timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for rep in 1 2; do
for n in "$@"; do
  export LZGPU_LIB=$PWD/lzma_b200/ab/lib_$n.so
  timeout 600 python bench.py --configs 3 --no-cpu-baseline --steps 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$n rep $rep: device', round(d['ms_per_step'],2), 'ms', round(d['value'],3), 'GB/s; e2e', round(d['e2e']['ms_per_step'],2), 'ms', round(d['e2e']['value'],3), d['e2e']['rank0_breakdown_ms'], 'config3 e2e', round(d['config3']['e2e']['ms'],2))"
  if [ $rep = 1 ]; then timeout 600 python scripts/bench_corpora.py --shapes text:148,text:2072 2>&1 | grep -v Warning; fi
done
done
