#!/bin/bash
mkdir -p gpurun_out
for eng in simt tcgen05_f32; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/train_ddp_check.py $eng 2>&1 | grep -E "^\{|Error|error" | tail -5
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --config 4 --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_train_bench_n2.json 2> gpurun_out/r2_train_bench_n2.err; echo "train bench n2 rc=$?"; tail -3 gpurun_out/r2_train_bench_n2.err; cut -c1-2400 gpurun_out/r2_train_bench_n2.json
python bench.py --config 4 --steps 10 --warmup 3 --no-gpu-eager-baseline > gpurun_out/r2_train_bench_n1.json 2>/dev/null; python - <<'PY'
import json
for f in ("gpurun_out/r2_train_bench_n1.json","gpurun_out/r2_train_bench_n2.json"):
    try:
        d=json.loads(open(f).read().strip().split("\n")[-1]); print(f, d["value"], d["ms_per_step"], d["phases_ms"], d["collective"])
    except Exception as e: print(f, "ERR", e)
PY
