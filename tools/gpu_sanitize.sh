#!/bin/bash
# memcheck + racecheck of the kernels rewritten this session on small shapes
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests -m gpu -q --tb=line -p no:cacheprovider -x -k "ffn or rows_resid or layernorm_loader or peaky or dwconv" > gpurun_out/sanitize_mem.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|error" gpurun_out/sanitize_mem.log | head -12
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests -m gpu -q --tb=line -p no:cacheprovider -x -k "ffn_fused or dwconv_bn or (attention and freq)" > gpurun_out/sanitize_race.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/sanitize_race.log | head -12
