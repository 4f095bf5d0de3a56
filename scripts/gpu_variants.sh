#!/bin/bash
# compare decoder tuning variants on the bench workload
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
for v in ${VARIANTS:-0 1 2 3}; do
  LZGPU_VARIANT=$v python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_v$v.json 2> gpurun_out/bench_v$v.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_v$v.json")); print("variant $v: %.2f ms/step  %.3f GB/s" % (d["ms_per_step"], d["value"]))
except Exception as e: print("variant $v failed", e, open("gpurun_out/bench_v$v.err").read()[-500:])
PY
done
if [ -n "$NCU_VARIANT" ]; then
export LZGPU_VARIANT=$NCU_VARIANT
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:lzgpu_decode -s 3 -c 1 -o gpurun_out/prof_v$NCU_VARIANT -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
fi
