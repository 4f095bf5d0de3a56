#!/bin/bash
# quick GPU iteration: parity tests, bench, one full ncu capture of the decode kernel
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -2 gpurun_out/bench.err; cat gpurun_out/bench.json
if [ "$1" != "noncu" ]; then
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:lzgpu_decode -s 3 -c 1 -o gpurun_out/prof -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
fi
