#!/bin/bash
# A/B of decoder tuning variants: parity of each variant, then lone-warp latency (148 units) and the bench shape (1024)
mkdir -p gpurun_out
for v in ${VARIANTS:-1 33}; do
  LZGPU_VARIANT=$v timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "not variants_agree" > gpurun_out/pytest_v$v.log 2>&1; echo "variant $v pytest rc=$?"; tail -3 gpurun_out/pytest_v$v.log
  echo "variant $v"; LZGPU_VARIANT=$v timeout 600 python scripts/bench_corpora.py --quick 2>&1 | tail -4
done
