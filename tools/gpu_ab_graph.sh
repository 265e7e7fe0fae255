#!/bin/bash
for i in 1 2; do
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('eager : ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), d['clocks']['sm_mhz'])"
timeout 600 python bench.py --steps 5 --warmup 3 --graphs --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('graphs: ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), d['clocks']['sm_mhz'])"
done
