#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_e2e.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err
echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_tc.json').read().strip().splitlines()[-1])
print('value',round(d['value'],1),'ms/step',round(d['ms_per_step'],1),'e2e',round(d['e2e']['value'],1))
for k,v in list(d['kernel_shares'].items())[:12]: print('   ',k,v, round(v*d['ms_per_step'],1),'ms')
PY
