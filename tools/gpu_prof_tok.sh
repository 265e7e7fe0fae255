#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"tok_gemm|gemm_tc_kernel|dwconv" -s 46 -c 6 -f -o /tmp/prof_tok python tools/profile_step.py > gpurun_out/prof_tok.log 2>&1
echo "rc=$?"
ncu -i /tmp/prof_tok.ncu-rep --page raw --csv > gpurun_out/prof_tok_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/prof_tok_raw.csv
