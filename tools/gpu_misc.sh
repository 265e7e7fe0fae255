#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 --batch 1 --clip-seconds 2 --graphs --no-cpu-baseline 2>gpurun_out/g.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('graphs B=1 2s: ms/step', round(d['ms_per_step'],3), 'launches', d['gpu_launches'])"
tail -3 gpurun_out/g.err
timeout 600 python bench.py --steps 5 --warmup 3 --graphs --no-cpu-baseline 2>gpurun_out/g2.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('graphs cfg2: ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1))"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -c 500
