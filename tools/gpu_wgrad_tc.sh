#!/bin/bash
# tcgen05 weight-gradient kernel: parity tests (both engines' callers), then the training profile / bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -q -p no:cacheprovider -x -k "wgrad or conv_forward or subpixel" 2>&1 | grep -E "passed|failed|Error|assert|rel_max" | cut -c1-300 | tail -12
timeout 600 python tools/train_prof.py tcgen05_f32 2>&1 | grep -E "==|wgrad" | head -14
