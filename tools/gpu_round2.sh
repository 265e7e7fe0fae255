#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "not tcgen05" --tb=short -p no:cacheprovider > gpurun_out/pytest_kernels_simt.log 2>&1
echo "simt kernels rc=$?" >> gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "tcgen05" --tb=short -p no:cacheprovider > gpurun_out/pytest_kernels_tc.log 2>&1
echo "tc kernels rc=$?" >> gpurun_out/summary.txt
export CUDA_LAUNCH_BLOCKING=1
for cfg in "simt 1" "simt 0" "tcgen05 0"; do
  timeout 600 python tools/gpu_report.py $cfg > gpurun_out/report_$(echo $cfg | tr ' ' '_').log 2>&1
  echo "report $cfg rc=$?" >> gpurun_out/summary.txt
done
unset CUDA_LAUNCH_BLOCKING
timeout 1200 python -m pytest tests/test_gpu_e2e.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_e2e.log 2>&1
echo "e2e rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
