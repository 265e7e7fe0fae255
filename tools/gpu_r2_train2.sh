#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_train.py -q -x -s -k "attention_train or training_step" 2>&1 | grep -E "passed|failed|Error|assert|^\[" | cut -c1-300 | tail -30
for eng in tcgen05_f32 simt; do
  python bench.py --config 4 --engine $eng --steps 10 --warmup 3 > gpurun_out/r2_train_bench_$eng.json 2> gpurun_out/r2_train_bench_$eng.err; echo "bench $eng rc=$?"; tail -3 gpurun_out/r2_train_bench_$eng.err; cut -c1-2500 gpurun_out/r2_train_bench_$eng.json
done
