#!/bin/bash
# 2-GPU box: the multi-GPU context paths of the C ABI (tests) and the strong / weak scaling bench at N = 2
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "all_visible_gpus or output_gaps or mixed_kinds or full_size or pinned" > gpurun_out/pytest_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_2gpu.log; tail -5 gpurun_out/pytest_2gpu.log
timeout 300 python scripts/bench_multi_ctx.py > gpurun_out/r02_one_process_2gpu.json 2> gpurun_out/mc.err; tail -2 gpurun_out/r02_one_process_2gpu.json
SECONDS=0
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$? wall=${SECONDS}s"
tail -5 gpurun_out/bench_n2.err; python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/r02_bench_n2.json") if l.startswith("{")][-1])
    print("N=2 headline", d["value"], "e2e", d["e2e"]["value"], "config3", d["config3"]["value"], d["config3"]["units_per_gpu"], "config5", d["config5"]["value"], d["config5"]["ms"], d["config5"]["units_per_gpu"])
except Exception as e:
    print("parse failed", e)
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02_bench_ref_n2.json 2>/dev/null; cut -c1-300 gpurun_out/r02_bench_ref_n2.json
