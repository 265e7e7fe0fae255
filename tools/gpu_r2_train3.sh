#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_train.py -q -s -k "attention_train or wgrad or conv_forward or subpixel or training_step" 2>&1 | grep -E "passed|failed|Error|assert|^\[" | cut -c1-300 | tail -30
python tools/train_prof.py ${1:-tcgen05_f32} 2>&1 | grep -E "==|attention|wgrad" 
