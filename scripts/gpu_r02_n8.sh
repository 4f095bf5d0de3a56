#!/bin/bash
# N-GPU box: bench.py under torchrun exactly as the driver launches it (weak-scaling headline, strong-scaling configs 3 and 5)
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi -L | wc -l; nproc; free -g | head -2
SECONDS=0
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n$N rc=$? wall=${SECONDS}s"
tail -4 gpurun_out/bench_n$N.err; python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/r02_bench_n$N.json") if l.startswith("{")][-1])
    print("N=$N headline", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), "config3", round(d["config3"]["value"],2), d["config3"]["units_per_gpu"], "config5", round(d["config5"]["value"],2), round(d["config5"]["ms"],1), d["config5"]["units_per_gpu"], "cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
except Exception as e:
    print("parse failed", e)
PY
