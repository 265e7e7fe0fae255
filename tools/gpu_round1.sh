#!/bin/bash
# first GPU session: kernel tests (SIMT first, tensor path in its own process), stage reports
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "not tcgen05" --tb=short -p no:cacheprovider > gpurun_out/pytest_kernels_simt.log 2>&1
echo "simt kernels rc=$?" >> gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "tcgen05" --tb=short -p no:cacheprovider > gpurun_out/pytest_kernels_tc.log 2>&1
echo "tc kernels rc=$?" >> gpurun_out/summary.txt
timeout 600 python tools/gpu_report.py simt 1 > gpurun_out/report_simt_1.log 2>&1
echo "report simt/1 rc=$?" >> gpurun_out/summary.txt
timeout 600 python tools/gpu_report.py simt 0 > gpurun_out/report_simt_0.log 2>&1
echo "report simt/0 rc=$?" >> gpurun_out/summary.txt
timeout 600 python tools/gpu_report.py tcgen05 0 > gpurun_out/report_tc_0.log 2>&1
echo "report tc/0 rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -5 gpurun_out/pytest_kernels_simt.log gpurun_out/pytest_kernels_tc.log
