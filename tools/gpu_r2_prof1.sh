#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train.csv python tools/train_one_step.py tcgen05_f32 0 > gpurun_out/launches_train.out 2>&1
echo "ncu train rc=$?"; wc -l gpurun_out/launches_train.csv
python tools/launch_shares.py gpurun_out/launches_train.csv "one generator training step (train-mode forward + backward), 4 x 2 s, tcgen05_f32 engine" 2>&1 | tee gpurun_out/launch_shares_train.txt | head -60
