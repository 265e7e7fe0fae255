#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -p no:cacheprovider -x -k attention > gpurun_out/pytest_attn.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest_attn.log
timeout 200 python tools/attn_tc_check.py big 2>&1 | grep -E "variant 3" 
