#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2b_bench.json").read().strip().split("\n")[-1])
print(d["value"], d["ms_per_step"], d["clocks"]); print({k:v["ms"] for k,v in d["kernel_rooflines"].items()})
PY
python -c "
import __graft_entry__ as g; g.smoke()"
python tools/train_prof.py tcgen05_f32 2>&1 | grep -E "==|attention|wgrad"
