#!/bin/bash
python -m pytest tests/test_gpu_train.py -q -s -k "training_step or gemm" 2>&1 | grep -E "passed|failed|Error|assert|^\[" | cut -c1-330 | tail -12
python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python bench.py --config 4 --steps 10 --warmup 3 --no-gpu-eager-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print(round(d['ms_per_step'],2), d['phases_ms'])"
