#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_fullsize.py -m gpu -q --tb=short -p no:cacheprovider --durations=5 > gpurun_out/pytest_fullsize.log 2>&1
echo "pytest rc=$?"; tail -n 30 gpurun_out/pytest_fullsize.log
