#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -k "conv or predict_matches or stages or rows_are_independent" > gpurun_out/pytest_k.log 2>&1
echo "pytest rc=$?"; tail -n 4 gpurun_out/pytest_k.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests -m gpu -q --tb=line -p no:cacheprovider -x -k "presplit" > gpurun_out/sanitize4_race.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/sanitize4_race.log | head -5
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_iter.json 2> gpurun_out/bench_iter.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_iter.json').read().strip().splitlines()[-1])
ms=d['ms_per_step']
print('ms/step', round(ms,2), 'audio-s/s', round(d['value'],1), 'clk', d['clocks'])
print({k: round(v*ms,2) for k,v in d['kernel_shares'].items() if v*ms>0.3})
PY
