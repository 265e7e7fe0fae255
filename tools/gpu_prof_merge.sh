#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"tok_gemm" -s 4 -c 2 -f -o /tmp/prof_merge python tools/merge_prof.py > gpurun_out/prof_merge.log 2>&1
echo "rc=$?"; tail -n 3 gpurun_out/prof_merge.log
ncu -i /tmp/prof_merge.ncu-rep --page raw --csv > gpurun_out/prof_merge_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/prof_merge_raw.csv | tee gpurun_out/ncu_full_merge.txt
