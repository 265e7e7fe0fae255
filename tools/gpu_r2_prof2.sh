#!/bin/bash
# ncu --set full of the training step's top kernels: attention_bwd_q / _k (time axis = first launches of a backward are the freq axis of TSCB_4; take 4), wgrad conv
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"attention_bwd_q|attention_bwd_k|attention_train_fwd" -c 6 -f -o /tmp/prof_attn_train python tools/train_one_step.py tcgen05_f32 0 > gpurun_out/prof_attn_train.log 2>&1
echo "rc=$?"
ncu -i /tmp/prof_attn_train.ncu-rep --page raw --csv > gpurun_out/prof_attn_train_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/prof_attn_train_raw.csv | tee gpurun_out/ncu_full_attention_train.txt
ncu -i /tmp/prof_attn_train.ncu-rep --page source --csv > gpurun_out/prof_attn_train_src.csv 2>/dev/null
python tools/ncu_source_top.py gpurun_out/prof_attn_train_src.csv 2>&1 | head -60 > gpurun_out/prof_attn_train_top.txt
ncu --set full --clock-control none --import-source on -k regex:"wgrad_kernel" -s 20 -c 3 -f -o /tmp/prof_wgrad python tools/train_one_step.py tcgen05_f32 0 > gpurun_out/prof_wgrad.log 2>&1
ncu -i /tmp/prof_wgrad.ncu-rep --page raw --csv > gpurun_out/prof_wgrad_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/prof_wgrad_raw.csv | tee gpurun_out/ncu_full_wgrad.txt
ncu -i /tmp/prof_wgrad.ncu-rep --page source --csv > gpurun_out/prof_wgrad_src.csv 2>/dev/null
python tools/ncu_source_top.py gpurun_out/prof_wgrad_src.csv 2>&1 | head -50 > gpurun_out/prof_wgrad_top.txt
