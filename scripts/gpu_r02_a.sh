#!/bin/bash
# round 2, first GPU pass: tests, smoke, bench with all BASELINE configurations
mkdir -p gpurun_out
{ nproc; free -g | head -2; nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv; } > gpurun_out/box.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
/usr/bin/time -v timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -25 gpurun_out/bench.err; cat gpurun_out/bench.json
