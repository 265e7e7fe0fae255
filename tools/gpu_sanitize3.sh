#!/bin/bash
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests -m gpu -q --tb=line -p no:cacheprovider -x -k "ffn_fused or backward or chain or shortest or default_init" > gpurun_out/sanitize3_mem.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid" gpurun_out/sanitize3_mem.log | head -8
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests -m gpu -q --tb=line -p no:cacheprovider -x -k "ffn_fused" > gpurun_out/sanitize3_race.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/sanitize3_race.log | head -8
