#!/bin/bash
# launch list + ncu --set full summaries of the top kernels for profiles/
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 135 -c 135 --csv --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/prof_launch.log 2>&1
echo "launch list rc=$?"
for spec in "attention_f16 8 2" "conv_split_tc 15 5" "ffn_fused 16 2" "gemm_tc_kernel 32 8" "dwconv 8 1" "inorm 31 4"; do
  set -- $spec
  ncu --set full --clock-control none -k regex:$1 -s $2 -c $3 -f -o /tmp/p_$1 python tools/profile_step.py > gpurun_out/prof_$1.log 2>&1
  ncu -i /tmp/p_$1.ncu-rep --page raw --csv > gpurun_out/raw_$1.csv 2>/dev/null
  echo "$1 rc=$?"
done
du -sh gpurun_out
