#!/bin/bash
# training step at N = 8 only (weak scaling, 4 x 2 s per GPU)
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29548 bench.py --config 4 --gpus 8 --steps 20 --warmup 5 > gpurun_out/train_final_n8.json 2> gpurun_out/train_final_n8.err
echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/train_final_n8.json").read().strip().split("\n")[-1]); print(round(d["value"],1), round(d["ms_per_step"],2), d["phases_ms"], d["collective"])
PY
