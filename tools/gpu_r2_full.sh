#!/bin/bash
# round-2 measurement set: full GPU tests, smoke, bench lines (configs[1] + reference arm + configs[2] + configs[4] training), profiles/r2 naming
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2f_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -n 4 gpurun_out/r2f_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/r2f_smoke.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/r2f_ref.json 2> gpurun_out/r2f_ref.err; echo "ref rc=$?"
timeout 600 python bench.py --config 2 --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline > gpurun_out/r2f_cfg2.json 2> gpurun_out/r2f_cfg2.err; echo "cfg2 rc=$?"
timeout 600 python bench.py --config 4 --steps 10 --warmup 3 > gpurun_out/r2f_cfg4.json 2> gpurun_out/r2f_cfg4.err; echo "cfg4 rc=$?"; tail -n 2 gpurun_out/r2f_cfg4.err
python - <<'PY'
import json
for f in ("r2f_bench","r2f_ref","r2f_cfg2","r2f_cfg4"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().split("\n")[-1])
        print(f, round(d["value"],2), round(d["ms_per_step"],2), d["config"]["workload"][:80], "e2e", d.get("e2e",{}).get("value"), d.get("phases_ms"), d.get("roofline",{}).get("frac"))
    except Exception as e: print(f,"ERR",e)
PY
