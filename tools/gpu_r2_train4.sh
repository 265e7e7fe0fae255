#!/bin/bash
# training kernels after a change: their parity tests, the per-kernel profile of one step, the bench line
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_train.py -q -p no:cacheprovider 2>&1 | grep -E "passed|failed|Error|assert" | cut -c1-300 | tail -12
timeout 600 python tools/train_prof.py tcgen05_f32 2>&1 | grep -E "==|attention|wgrad" | head -30
timeout 600 python bench.py --config 4 --steps 20 --warmup 5 --no-gpu-eager-baseline > gpurun_out/train_bench_n1_s4.json 2>/dev/null
python - <<'PY'
import json
d=json.loads(open("gpurun_out/train_bench_n1_s4.json").read().strip().split("\n")[-1]); print(round(d["value"],1), round(d["ms_per_step"],2), d["phases_ms"])
PY
