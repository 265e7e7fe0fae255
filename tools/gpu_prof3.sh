#!/bin/bash
mkdir -p gpurun_out
K=${1:-ffn_fused}
SKIP=${2:-16}
ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c 1 -f -o /tmp/prof_k python tools/profile_step.py > gpurun_out/prof_k.log 2>&1
echo "rc=$?"
ncu -i /tmp/prof_k.ncu-rep --page raw --csv > gpurun_out/prof_${K}_raw.csv 2>/dev/null
ncu -i /tmp/prof_k.ncu-rep --page source --csv > gpurun_out/prof_${K}_source.csv 2>/dev/null
ncu -i /tmp/prof_k.ncu-rep --page details > gpurun_out/prof_${K}_details.txt 2>/dev/null
