#!/bin/bash
# ncu --set full of the tcgen05 attention kernel (variant 3): time axis
mkdir -p gpurun_out
PROF_B=${PROF_B:-64} ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 2 -c 1 -f -o /tmp/prof_tc python tools/attn_prof.py 3 ${1:-time} > gpurun_out/prof_tc.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/prof_tc.log
ncu -i /tmp/prof_tc.ncu-rep --page raw --csv > gpurun_out/prof_tc_raw.csv 2>/dev/null
ncu -i /tmp/prof_tc.ncu-rep --page source --csv > gpurun_out/prof_tc_source.csv 2>/dev/null
ncu -i /tmp/prof_tc.ncu-rep --page details > gpurun_out/prof_tc_details.txt 2>/dev/null
python tools/ncu_summary.py gpurun_out/prof_tc_raw.csv
python tools/ncu_source_top.py gpurun_out/prof_tc_source.csv 60
