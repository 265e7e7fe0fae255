#!/bin/bash
# round 2, call A: full GPU suite + default bench line (with the eager-GPU leg) + sanitizer passes the round-1 verdict asked for
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"; cut -c1-900 gpurun_out/r2a_bench.json
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests -m gpu -q --tb=line -p no:cacheprovider -k "attention_peaky or rows_resid or layernorm_loader or presplit" > gpurun_out/r2a_race.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/r2a_race.log | head -8
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests -m gpu -q --tb=line -p no:cacheprovider -k "attention_peaky or rows_resid or layernorm_loader or glu" > gpurun_out/r2a_mem.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid" gpurun_out/r2a_mem.log | head -8
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/sanitize_attn_tc3.py > gpurun_out/r2a_race_tc3.log 2>&1
echo "racecheck tc3 rc=$?"; grep -E "RACECHECK SUMMARY|ok|hazard" gpurun_out/r2a_race_tc3.log | head -12
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_attn_tc3.py > gpurun_out/r2a_mem_tc3.log 2>&1
echo "memcheck tc3 rc=$?"; grep -E "ERROR SUMMARY|ok|Invalid" gpurun_out/r2a_mem_tc3.log | head -12
