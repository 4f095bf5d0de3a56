#!/bin/bash
# round 2, final build: GPU tier, smoke, bench (both arms), ncu launch list + full capture, C++ reader
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/r02d_bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -2 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02d_bench_reference_arm.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
python bench.py --steps 2 --warmup 3 --configs '' --no-e2e --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r02d_launches.csv python bench.py --steps 2 --warmup 3 --configs '' --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lzgpu_sm_kernel -s 3 -c 1 -o gpurun_out/r02d_prof1024 -f python bench.py --steps 2 --warmup 3 --configs '' --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
tests/cpp/_build/reader_test tests/golden/ref_assets 2>&1 | tail -2
python - <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import bench
blob, offs, lens, _ = bench.build_corpus3(256, 1 << 20)
with open("/tmp/s.lzma2", "wb") as f:
    for i in range(4096):
        k = i % 256
        f.write(blob[offs[k]:offs[k] + lens[k]].tobytes())
    f.write(b"\0")
PY
tests/cpp/_build/reader2_bench /tmp/s.lzma2 1048576 1073741824 2>&1 | tail -1 | tee gpurun_out/r02d_reader2.jsonl
rm -f /tmp/s.lzma2
ls -la gpurun_out | grep r02d
