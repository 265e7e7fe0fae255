#!/bin/bash
# ncu of the tcgen05 weight-gradient kernel (conv layers 3 / 4 of a dense block + a token layer) and the launch list of one training step
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"wgrad_tc_kernel" -s 20 -c 4 -f -o /tmp/prof_wgrad_tc python tools/train_one_step.py tcgen05_f32 0 > gpurun_out/prof_wgrad_tc.log 2>&1
echo "rc=$?"
ncu -i /tmp/prof_wgrad_tc.ncu-rep --page raw --csv > gpurun_out/prof_wgrad_tc_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/prof_wgrad_tc_raw.csv | tee gpurun_out/ncu_full_wgrad_tc.txt
ncu -i /tmp/prof_wgrad_tc.ncu-rep --page source --csv > gpurun_out/prof_wgrad_tc_src.csv 2>/dev/null
python tools/ncu_source_top.py gpurun_out/prof_wgrad_tc_src.csv 2>&1 | head -45 | tee gpurun_out/prof_wgrad_tc_top.txt
