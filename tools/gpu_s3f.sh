#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "default rc=$?"; tail -c 300 gpurun_out/bench_default.err
timeout 900 python bench.py --steps 2 --warmup 3 --batch 16 --clip-seconds 30 --no-cpu-baseline > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err
echo "cfg3 rc=$?"; tail -c 300 gpurun_out/bench_cfg3.err
python - <<'PY'
import json
for f in ("default","cfg3"):
    try:
        d=json.loads(open(f'gpurun_out/bench_{f}.json').read().strip().splitlines()[-1])
        print(f, 'audio-s/s', round(d['value'],1), 'ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), {k:v for k,v in list(d.get('kernel_shares',{}).items())[:5]}, d.get('clocks'))
    except Exception as e: print(f, 'failed', e)
PY
