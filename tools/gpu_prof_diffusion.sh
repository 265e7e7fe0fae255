#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_diffusion.csv python tools/profile_diffusion_step.py > gpurun_out/prof_diffusion.log 2>&1
echo "rc=$?"
python tools/launch_shares.py gpurun_out/launches_diffusion.csv > gpurun_out/launch_shares_diffusion.txt; head -n 16 gpurun_out/launch_shares_diffusion.txt
