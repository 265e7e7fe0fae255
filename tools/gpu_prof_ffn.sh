#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:ffn_fused -s 16 -c 1 -f -o /tmp/prof_ffn python tools/profile_step.py > gpurun_out/prof_ffn.log 2>&1
echo "ffn rc=$?"
ncu -i /tmp/prof_ffn.ncu-rep --page raw --csv > gpurun_out/prof_ffn_raw.csv 2>/dev/null
ncu -i /tmp/prof_ffn.ncu-rep --page source --csv > gpurun_out/prof_ffn_source.csv 2>/dev/null
