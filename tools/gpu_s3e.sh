#!/bin/bash
mkdir -p gpurun_out
export SEB200_T5_MODE=${1:-6}
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -p no:cacheprovider -x -k "attention and 3" > gpurun_out/pytest_attn.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest_attn.log
timeout 200 python tools/attn_tc_check.py big 2>&1 | grep -E "variant 3" 
