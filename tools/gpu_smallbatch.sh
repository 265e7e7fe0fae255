#!/bin/bash
# does a micro-batch whose activations fit the 126 MB L2 run faster per clip than the 64-clip batch? (CUDA-graph replay per micro-batch)
mkdir -p gpurun_out
for b in 1 2 4 8 64; do
  g="--graphs"; [ $b -eq 64 ] && g=""
  timeout 300 python bench.py --batch $b $g --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_b$b.json 2> gpurun_out/bench_b$b.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_b$b.json').read().strip().splitlines()[-1])
print('batch $b', 'ms/step', round(d['ms_per_step'],3), 'audio-s/s', round(d['value'],1), 'ms/clip', round(d['ms_per_step']/$b,3), d['clocks']['sm_mhz'])
PY
done
