#!/bin/bash
# two GPUs: training step with the discriminator under DistributedDataParallel vs one coalesced all-reduce of its gradients
mkdir -p gpurun_out
for mode in "" "--disc-ddp"; do
  tag=${mode:+ddp}; tag=${tag:-flat}
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --config 4 --gpus 2 --steps 20 --warmup 5 $mode > gpurun_out/train_n2_disc_$tag.json 2> gpurun_out/train_n2_disc_$tag.err
  echo "rc=$? $tag"; tail -3 gpurun_out/train_n2_disc_$tag.err | cut -c1-300
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/train_n2_disc_*.json")):
    try:
        d=json.loads(open(f).read().strip().split("\n")[-1]); print(f, round(d["value"],1), round(d["ms_per_step"],2), d["phases_ms"], d["collective"])
    except Exception as e: print(f, "ERR", e)
PY
