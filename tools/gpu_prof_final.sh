#!/bin/bash
# launch list + ncu --set full summaries of the top kernels for profiles/ (pass 1 of tools/profile_step.py warms up, pass 2 is profiled)
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 135 -c 135 --csv --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/prof_launch.log 2>&1
echo "launch list rc=$?"
# name, launches to skip (= launches of that kernel in the warm-up pass), launches to capture
for spec in "attention 8 2" "conv_y3 12 4" "ffn_fused 16 2" "tok_gemm 32 4" "dwconv 8 2" "inorm 31 4"; do
  set -- $spec
  ncu --set full --clock-control none -k regex:$1 -s $2 -c $3 -f -o /tmp/p_$1 python tools/profile_step.py > gpurun_out/prof_$1.log 2>&1
  ncu -i /tmp/p_$1.ncu-rep --page raw --csv > gpurun_out/raw_$1.csv 2>/dev/null
  echo "$1 rc=$?"
  python tools/ncu_summary.py gpurun_out/raw_$1.csv > gpurun_out/ncu_full_$1.txt 2>/dev/null
  cat gpurun_out/ncu_full_$1.txt | cut -c1-400
done
python tools/launch_shares.py gpurun_out/launches.csv > gpurun_out/launch_shares.txt 2>/dev/null; head -30 gpurun_out/launch_shares.txt
du -sh gpurun_out
