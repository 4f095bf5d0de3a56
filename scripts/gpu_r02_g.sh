#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --configs 3,4 --steps 5 > gpurun_out/bench_g.json 2> gpurun_out/bench_g.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_g.json"))
print("headline", d["ms_per_step"], d["value"], "e2e", d["e2e"]["ms_per_step"], d["e2e"]["value"], d["e2e"]["rank0_breakdown_ms"])
print("config3", d["config3"]["ms"], d["config3"]["e2e"]["ms"], "config4", d["config4"]["ms"], d["config4"]["kernel_ms"], d["config4"]["units_total"])
PY
timeout 300 python scripts/bench_corpora.py --quick-random 2>/dev/null
