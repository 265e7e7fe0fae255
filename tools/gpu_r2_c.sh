#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2c_bench.err
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2c_ref.json 2> gpurun_out/r2c_ref.err; echo "ref rc=$?"
python bench.py --config 2 --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline > gpurun_out/r2c_cfg2.json 2> gpurun_out/r2c_cfg2.err; echo "cfg2 rc=$?"
python - <<'PY'
import json
for f in ("r2c_bench","r2c_ref","r2c_cfg2"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().split("\n")[-1])
        print(f, round(d["value"],2), round(d["ms_per_step"],2), d["config"]["workload"][:90], d.get("e2e",{}).get("value"), d.get("cpu_baseline"), d.get("gpu_eager_baseline"), d.get("roofline",{}).get("frac"))
    except Exception as e: print(f,"ERR",e)
PY
