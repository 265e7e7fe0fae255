#!/bin/bash
# round 2: ncu launch list of one inference step at configs[1] + --set full of the two kernels whose traffic changed (fp16 v)
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 140 -c 140 --csv --log-file gpurun_out/launches_r2.csv python tools/profile_step.py > gpurun_out/launches_r2.out 2>&1
echo "launch list rc=$?"
python tools/launch_shares.py gpurun_out/launches_r2.csv "one forward pass of the hot path at BASELINE configs[1] (64 x 4 s), round 2 (v in fp16)" | tee gpurun_out/launch_shares_r2.txt | head -30
ncu --set full --clock-control none --import-source on -k regex:"dwconv_bn_swish|tok_gemm_kernel<64, 2" -s 4 -c 4 -f -o /tmp/prof_v16 python tools/profile_step.py > gpurun_out/prof_v16.log 2>&1
ncu -i /tmp/prof_v16.ncu-rep --page raw --csv > gpurun_out/prof_v16_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/prof_v16_raw.csv | tee gpurun_out/ncu_full_v16.txt
