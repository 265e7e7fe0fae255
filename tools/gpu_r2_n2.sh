#!/bin/bash
mkdir -p gpurun_out
for eng in simt tcgen05_f32; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/train_ddp_check.py $eng 2>&1 | grep -E "^\{|Error|error" | tail -3
done
for eng in simt tcgen05_f32; do
python bench.py --config 4 --engine $eng --steps 10 --warmup 3 --no-gpu-eager-baseline > gpurun_out/r2_train_bench_n1_$eng.json 2>/dev/null
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --config 4 --engine $eng --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_train_bench_n2_$eng.json 2> gpurun_out/r2_train_bench_n2.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_train_bench_n?_*.json")):
    try:
        d=json.loads(open(f).read().strip().split("\n")[-1]); print(f, round(d["value"],1), round(d["ms_per_step"],2), d["phases_ms"])
    except Exception as e: print(f, "ERR", e)
PY
