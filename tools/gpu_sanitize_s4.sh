#!/bin/bash
# compute-sanitizer over the kernels changed in round 2, session 4: tcgen05 weight gradients + parallel finish kernels, fringe-aware training attention
mkdir -p gpurun_out
K='wgrad or conv_forward_dgrad or subpixel_conv_backward or attention_train or layernorm_backward or batchnorm_train or depthwise_conv_train'
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_train.py -q --tb=line -p no:cacheprovider -x -k "$K" > gpurun_out/sanitize_s4_mem.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid" gpurun_out/sanitize_s4_mem.log | head -8
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_train.py -q --tb=line -p no:cacheprovider -x -k "attention_train or wgrad_rows or layernorm_backward or batchnorm_train" > gpurun_out/sanitize_s4_race.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard|Race reported" gpurun_out/sanitize_s4_race.log | head -12
