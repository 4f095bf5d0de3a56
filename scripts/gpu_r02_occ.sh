#!/bin/bash
# config 5's per-GPU shape at N = 8 (2048 units, one wave) under an occupancy cap: does the slowest unit finish sooner
# when fewer units share its SM?
mkdir -p gpurun_out
for k in 0 13 12 11 10 9 8 7; do
  LZGPU_MAX_CTAS_PER_SM=$k timeout 300 python bench.py --configs 5 --c5-units 2048 --no-e2e --no-cpu-baseline --steps 3 > gpurun_out/occ_$k.json 2> gpurun_out/occ_$k.err
  python -c "
import json
d=json.load(open('gpurun_out/occ_$k.json')); c=d['config5']; print('max CTAs/SM $k: config5(2048 units)', round(c['ms'],1), 'ms', round(c['value'],3), 'GB/s   headline', round(d['ms_per_step'],2), 'ms')"
done
