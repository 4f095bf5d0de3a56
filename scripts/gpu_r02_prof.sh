#!/bin/bash
# round 2, scheduler kernel: ncu launch list of the bench command and one full capture of the dominant kernel
# (bench shape: 1024 units), plus one of the single-wave shape (2072 units), after the plain command exited 0
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --configs '' --no-e2e --no-cpu-baseline > gpurun_out/plain.log 2>&1 || { echo "plain bench failed"; tail -5 gpurun_out/plain.log; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r02b_launches.csv python bench.py --steps 2 --warmup 3 --configs '' --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lzgpu_sm_kernel -s 3 -c 1 -o gpurun_out/r02b_prof1024 -f python bench.py --steps 2 --warmup 3 --configs '' --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lzgpu_sm_kernel -s 2 -c 1 -o gpurun_out/r02b_prof2072 -f python scripts/bench_corpora.py --shapes text:2072 > gpurun_out/ncu_full2.log 2>&1; echo "ncu full (2072) rc=$?"; tail -2 gpurun_out/ncu_full2.log
ls -la gpurun_out | grep r02b
