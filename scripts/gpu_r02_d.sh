#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 900 python scripts/bench_reader2.py --gib 4 > gpurun_out/r02_reader2.jsonl 2> gpurun_out/reader2.err; echo "reader2 rc=$?"; cat gpurun_out/r02_reader2.jsonl; tail -3 gpurun_out/reader2.err
