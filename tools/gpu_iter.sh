#!/bin/bash
# quick iteration: selected GPU tests + bench (no CPU leg); usage: tools/gpu_iter.sh "<pytest -k expr>"
mkdir -p gpurun_out
K="${1:-}"
if [ -n "$K" ]; then
  timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -k "$K" > gpurun_out/pytest_iter.log 2>&1
else
  timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_iter.log 2>&1
fi
echo "pytest rc=$?"; tail -15 gpurun_out/pytest_iter.log
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_iter.json 2> gpurun_out/bench_iter.err
echo "bench rc=$?"; tail -3 gpurun_out/bench_iter.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_iter.json').read().strip().splitlines()[-1])
ms=d['ms_per_step']
print('ms/step', round(ms,2), 'audio-s/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'clk', d['clocks'])
print({k: round(v*ms,2) for k,v in d['kernel_shares'].items() if v*ms>0.3})
PY
