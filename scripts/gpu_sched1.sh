#!/bin/bash
# SM-resident scheduler, first GPU pass: the whole GPU tier on the new kernel (hangs bounded by timeout), then A/B of
# the scheduler against the one-warp CTAs (LZGPU_SCHED=0) on the shapes where the sub-partition balance matters.
mkdir -p gpurun_out
export LZGPU_LIB=$PWD/lzma_b200/ab/lib_sched.so
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_sched.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_sched.log
for sch in 1 0; do
  echo "== LZGPU_SCHED=$sch"
  LZGPU_SCHED=$sch timeout 600 python scripts/bench_corpora.py --shapes text:148,text:1024,text:1924,text:2072,mixed:1024,mixed:2072 2>&1 | grep -v Warning
  LZGPU_SCHED=$sch timeout 600 python bench.py --configs 5 --c5-units 2048 --no-e2e --no-cpu-baseline --steps 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=d['config5']; print('config5 2048 units:', round(c['ms'],1), 'ms', round(c['value'],3), 'GB/s', 'headline', d['ms_per_step'])"
done
