#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_e2e.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err
echo "bench rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -s 158 -c 158 --csv --log-file gpurun_out/launches_r1.csv python tools/profile_step.py > gpurun_out/prof_launch.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:attention_mma -s 8 -c 2 -f -o gpurun_out/prof_attn python tools/profile_step.py > gpurun_out/prof_attn.log 2>&1
echo "attn rc=$?"
ncu -i gpurun_out/prof_attn.ncu-rep --page raw --csv > gpurun_out/prof_attn_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_attn.ncu-rep --page source --csv --kernel-id :::1 > gpurun_out/prof_attn_source.csv 2>/dev/null
ncu --set full --clock-control none -k regex:gemm_tc -s 79 -c 24 -f -o /tmp/prof_gemm python tools/profile_step.py > gpurun_out/prof_gemm.log 2>&1
echo "gemm rc=$?"
ncu -i /tmp/prof_gemm.ncu-rep --page raw --csv > gpurun_out/prof_gemm_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:dwconv -s 8 -c 2 -f -o /tmp/prof_dw python tools/profile_step.py > gpurun_out/prof_dw.log 2>&1
ncu -i /tmp/prof_dw.ncu-rep --page raw --csv > gpurun_out/prof_dw_raw.csv 2>/dev/null
echo "dw rc=$?"
du -sh gpurun_out; ls -la gpurun_out | head -30
