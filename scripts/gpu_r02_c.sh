#!/bin/bash
# round 2, second GPU pass: regime sweep with the shipped kernel, per-GPU shapes of config 5 at N = 2/4/8, A/B of the
# launch bounds, ncu launch list + one full capture
mkdir -p gpurun_out
timeout 900 python scripts/bench_corpora.py > gpurun_out/r02_shapes.jsonl 2> gpurun_out/shapes.err; echo "shapes rc=$?"; cat gpurun_out/r02_shapes.jsonl
for n in 2048 4096 8192; do
  timeout 600 python bench.py --configs 5 --c5-units $n --no-e2e --no-cpu-baseline --steps 3 > gpurun_out/c5_$n.json 2> gpurun_out/c5_$n.err; echo "c5 $n rc=$?"
  python - <<PY
import json
d=json.load(open("gpurun_out/c5_$n.json")); c=d["config5"]; print("config5 per-GPU shape", c["units_total"], "units:", round(c["ms"],1), "ms", round(c["value"],3), "GB/s")
PY
done
for lib in liblzgpu.so liblzgpu_lb8.so; do
  LZGPU_LIB=$PWD/lzma_b200/$lib timeout 300 python bench.py --configs '' --no-e2e --no-cpu-baseline --steps 5 > gpurun_out/ab_$lib.json 2>/dev/null
  python -c "import json; d=json.load(open('gpurun_out/ab_$lib.json')); print('$lib', d['ms_per_step'], d['value'])"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --configs '' --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lzgpu_decode -s 3 -c 1 -o gpurun_out/r02_prof1024 -f python bench.py --steps 2 --warmup 3 --configs '' --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out | head -30
