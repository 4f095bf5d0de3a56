#!/bin/bash
# round 2, scheduler kernel: GPU tier, smoke, regime sweep, full bench line, reference arm
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python scripts/bench_corpora.py > gpurun_out/r02b_shapes.jsonl 2> gpurun_out/shapes.err; echo "shapes rc=$?"; cat gpurun_out/r02b_shapes.jsonl
timeout 900 python bench.py > gpurun_out/r02b_bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err; cat gpurun_out/r02b_bench.json
