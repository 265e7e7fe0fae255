#!/bin/bash
# the ncu launch list of the bench command itself (cold-cache, serialised: compare SHARES with the bench line's kernel_shares)
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
echo "rc=$?"
python tools/launch_shares.py gpurun_out/launches_bench.csv > gpurun_out/launch_shares_bench.txt; head -n 14 gpurun_out/launch_shares_bench.txt
gzip -f gpurun_out/launches_bench.csv; ls -la gpurun_out/launches_bench.csv.gz
