#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:attention_v4 -s 8 -c 2 -f -o /tmp/prof_attn python tools/profile_step.py > gpurun_out/prof_attn.log 2>&1
echo "attn rc=$?"
ncu -i /tmp/prof_attn.ncu-rep --page raw --csv > gpurun_out/prof_attn_raw.csv 2>/dev/null
ncu -i /tmp/prof_attn.ncu-rep --page source --csv > gpurun_out/prof_attn_source.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/prof_attn_raw.csv
