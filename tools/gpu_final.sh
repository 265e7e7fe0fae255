#!/bin/bash
# end-of-session measurement set: full GPU tests, default bench line (with the CPU leg), reference arm, configs[2] and configs[0]
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -n 3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "default rc=$?"
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
echo "reference rc=$?"
timeout 900 python bench.py --steps 2 --warmup 3 --batch 16 --clip-seconds 30 --no-cpu-baseline > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err
echo "cfg3 rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 --batch 1 --clip-seconds 2 --no-cpu-baseline > gpurun_out/bench_cfg1.json 2> gpurun_out/bench_cfg1.err
echo "cfg1 rc=$?"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 3 gpurun_out/smoke.log
python - <<'PY'
import json
for f in ("default","reference","cfg3","cfg1"):
    try:
        d=json.loads(open(f'gpurun_out/bench_{f}.json').read().strip().splitlines()[-1])
        print(f, 'value', round(d['value'],2), d['unit'], 'ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), d.get('clocks'), d.get('cpu_baseline'))
    except Exception as e: print(f, 'failed', e)
PY
