#!/bin/bash
# eight GPUs: where the training step's last 6 % go -- the metric-label pipeline's host work (A/B with fixed label tensors)
mkdir -p gpurun_out
python -c "import os; print('host cores', os.cpu_count())"
for mode in "--fixed-labels" ""; do
  tag=${mode:+fixed}; tag=${tag:-pipeline}
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29548 bench.py --config 4 --gpus 8 --steps 20 --warmup 5 $mode > gpurun_out/train_n8_$tag.json 2> gpurun_out/train_n8_$tag.err
  echo "rc=$? $tag"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/train_n8_*.json")):
    try:
        d=json.loads(open(f).read().strip().split("\n")[-1]); print(f, round(d["value"],1), round(d["ms_per_step"],2), d["phases_ms"])
    except Exception as e: print(f, "ERR", e)
PY
