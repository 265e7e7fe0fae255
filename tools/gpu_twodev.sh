#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_e2e.py -m gpu -q --tb=short -p no:cacheprovider -k "second_device or two_streams or predict_matches" > gpurun_out/pytest_twodev.log 2>&1
echo "pytest rc=$?"; tail -n 15 gpurun_out/pytest_twodev.log
