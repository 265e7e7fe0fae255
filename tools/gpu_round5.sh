#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_e2e.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
for mk in 1000 12 6 0; do
  SEB200_CONV_PERSIST_MAXK=$mk timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_mk$mk.json 2>/dev/null
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_mk$mk.json').read().strip().splitlines()[-1])
print('maxk=$mk ms/step',round(d['ms_per_step'],1), 'dconv', round(d['kernel_shares']['dconv']*d['ms_per_step'],1), 'conv2', round(d['kernel_shares']['conv2']*d['ms_per_step'],2))
PY
done
