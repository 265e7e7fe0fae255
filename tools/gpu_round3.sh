#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest gpu rc=$?" >> gpurun_out/summary.txt
timeout 300 python __graft_entry__.py > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/summary.txt
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err
echo "bench tc rc=$?" >> gpurun_out/summary.txt
timeout 1200 python bench.py --steps 2 --warmup 3 --engine simt --no-cpu-baseline > gpurun_out/bench_simt.json 2> gpurun_out/bench_simt.err
echo "bench simt rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -3 gpurun_out/pytest_gpu.log
cat gpurun_out/smoke.log | tail -2
