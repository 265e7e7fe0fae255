#!/bin/bash
# usage: tools/gpu_prof_k.sh <kernel regex> <skip> ; full ncu capture of one launch inside the real step + source-level stalls
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c 1 -f -o /tmp/prof_k python tools/profile_step.py > gpurun_out/prof_k.log 2>&1
echo "rc=$?"
ncu -i /tmp/prof_k.ncu-rep --page raw --csv > gpurun_out/prof_k_raw.csv 2>/dev/null
ncu -i /tmp/prof_k.ncu-rep --page source --csv > gpurun_out/prof_k_source.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/prof_k_raw.csv
