#!/bin/bash
# ncu --set full + source page of the hybrid attention kernel (variant 4) at a reduced batch (PROF_B), time axis n = 641
mkdir -p gpurun_out
export PROF_B=${PROF_B:-16}
ncu --set full --clock-control none --import-source on -k regex:attention_tc8 -s 2 -c 1 -f -o /tmp/prof_attn8 python tools/attn_prof.py 4 time > gpurun_out/prof_attn8.log 2>&1
echo "rc=$?"
ncu -i /tmp/prof_attn8.ncu-rep --page raw --csv > gpurun_out/prof_attn8_raw.csv 2>/dev/null
ncu -i /tmp/prof_attn8.ncu-rep --page source --csv > gpurun_out/prof_attn8_source.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/prof_attn8_raw.csv | head -60
python tools/ncu_source_top.py gpurun_out/prof_attn8_source.csv 2>&1 | head -40
