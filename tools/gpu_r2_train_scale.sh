#!/bin/bash
# training step (BASELINE configs[4]) at N = 1, 2, 4, 8 on one box: weak scaling, 4 x 2 s per GPU
mkdir -p gpurun_out
python bench.py --config 4 --steps 20 --warmup 5 --no-gpu-eager-baseline > gpurun_out/train_scale_n1.json 2> gpurun_out/train_scale_n1.err
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py --config 4 --gpus $n --steps 20 --warmup 5 > gpurun_out/train_scale_n$n.json 2> gpurun_out/train_scale_n$n.err
  echo "n=$n rc=$?"
done
python - <<'PY'
import json
base=None
for n in (1,2,4,8):
    try:
        d=json.loads(open(f"gpurun_out/train_scale_n{n}.json").read().strip().split("\n")[-1])
        if n==1: base=d["value"]
        print(n, round(d["value"],1), "audio-s/s", round(d["ms_per_step"],2), "ms/step  eff", round(d["value"]/(n*base),3), d["phases_ms"], d["collective"]["allreduce_ms"], d["clocks"])
    except Exception as e: print(n, "ERR", e)
PY
