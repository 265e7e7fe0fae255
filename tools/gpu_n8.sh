#!/bin/bash
# usage: tools/gpu_n8.sh N   -- weak-scaling bench (configs[1] per GPU) and the configs[3] fixed job on N GPUs of one box
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -n 8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "n$N rc=$?"; tail -c 300 gpurun_out/bench_n$N.err; tail -c 900 gpurun_out/bench_n$N.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --total-clips 4096 --steps 1 --warmup 1 > gpurun_out/bench_cfg4_n$N.json 2> gpurun_out/bench_cfg4_n$N.err
echo "cfg4 n$N rc=$?"; tail -c 300 gpurun_out/bench_cfg4_n$N.err; tail -c 900 gpurun_out/bench_cfg4_n$N.json
