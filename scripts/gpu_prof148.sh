#!/bin/bash
# ncu full capture of the decode kernel in the lone-warp shape (148 units), after a clean run
mkdir -p gpurun_out
V=${LZGPU_VARIANT:-33}
export LZGPU_VARIANT=$V
timeout 300 python scripts/bench_corpora.py --quick 2>&1 | tail -2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lzgpu_decode -s 3 -c 1 -o gpurun_out/prof148_v$V -f python scripts/bench_corpora.py --quick > gpurun_out/ncu148.log 2>&1; tail -1 gpurun_out/ncu148.log
