#!/bin/bash
# usage: tools/gpu_k.sh "<pytest -k expression>" [test file]
mkdir -p gpurun_out
timeout 900 python -m pytest ${2:-tests} -m gpu -q --tb=short -p no:cacheprovider -k "$1" > gpurun_out/pytest_k.log 2>&1
echo "pytest rc=$?"; tail -n 40 gpurun_out/pytest_k.log
