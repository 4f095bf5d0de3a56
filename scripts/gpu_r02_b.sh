#!/bin/bash
mkdir -p gpurun_out
{ nproc; free -g | head -2; nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv; } > gpurun_out/box.txt 2>&1
SECONDS=0
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$? wall=${SECONDS}s"
tail -25 gpurun_out/bench.err; cat gpurun_out/bench.json
SECONDS=0
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench2.json 2> gpurun_out/bench2.err; echo "bench (cached corpora) rc=$? wall=${SECONDS}s"
