#!/bin/bash
mkdir -p gpurun_out
LZGPU_LIB=$PWD/lzma_b200/ab/lib_$1.so timeout 900 python -m pytest tests -m gpu -x -q -k "pinned or gaps or reader or xz or folders or mixed or sm_scheduler" 2>&1 | tail -2
for rep in 1 2; do
for n in "$@"; do
  export LZGPU_LIB=$PWD/lzma_b200/ab/lib_$n.so
  timeout 600 python bench.py --configs 3 --no-cpu-baseline --steps 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$n rep $rep: device', round(d['ms_per_step'],2), 'ms', round(d['value'],3), 'GB/s; e2e', round(d['e2e']['ms_per_step'],2), 'ms', round(d['e2e']['value'],3), d['e2e']['rank0_breakdown_ms'], 'config3 e2e', round(d['config3']['e2e']['ms'],2))"
done
done
