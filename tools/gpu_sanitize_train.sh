#!/bin/bash
# compute-sanitizer over the training-step kernels (round 2): memcheck on every kernel-level training test, racecheck on the shared-memory-heavy ones
mkdir -p gpurun_out
K='device_packer or train_gemm or wgrad or conv_forward_dgrad or subpixel_conv_backward or elementwise_forward or dropout_mask or layernorm_backward or batchnorm_train or depthwise_conv_train or instancenorm_prelu_backward or head_conv or mask_tail or conv1x1_in3 or attention_train'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_train.py -q --tb=line -p no:cacheprovider -x -k "$K" > gpurun_out/sanitize_train_mem.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid" gpurun_out/sanitize_train_mem.log | head -8
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_train.py -q --tb=line -p no:cacheprovider -x -k "attention_train or wgrad_rows or layernorm_backward or batchnorm_train or instancenorm_prelu_backward" > gpurun_out/sanitize_train_race.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/sanitize_train_race.log | head -8
