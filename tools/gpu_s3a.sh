#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_e2e.py -m gpu -q --tb=short -p no:cacheprovider -k "batch_stft or golden" > gpurun_out/pytest_a18.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_a18.log
timeout 600 python bench.py --total-clips 512 --steps 2 --warmup 3 > gpurun_out/bench_cfg4_512.json 2> gpurun_out/bench_cfg4.err
echo "cfg4 rc=$?"; cat gpurun_out/bench_cfg4_512.json; tail -3 gpurun_out/bench_cfg4.err
