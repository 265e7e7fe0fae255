#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"tok_gemm" -s 32 -c 2 -f -o /tmp/prof_tok python tools/profile_step.py > gpurun_out/prof_tok.log 2>&1
echo "rc=$?"
ncu -i /tmp/prof_tok.ncu-rep --page source --csv > gpurun_out/prof_tok_source.csv 2>/dev/null
