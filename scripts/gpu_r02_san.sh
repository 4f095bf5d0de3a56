#!/bin/bash
mkdir -p gpurun_out
which compute-sanitizer; 
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "alone_cases or lzma2_uncompressed or long_and_overlapping or lzma2_cases" > gpurun_out/r02_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -15 gpurun_out/r02_sanitizer_memcheck.log
