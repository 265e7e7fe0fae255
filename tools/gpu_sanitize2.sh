#!/bin/bash
# memcheck + racecheck of the kernels added / changed in the fourth session (merge-block token GEMMs, diffusion kernels, packed-arithmetic FFN / LN warps)
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests -m gpu -q --tb=line -p no:cacheprovider -x -k "merge_block or diffusion_embed or ffn_fused or rows_resid or layernorm_loader or (diffusion_module and tcgen05) or (reverse_steps and tcgen05)" > gpurun_out/sanitize2_mem.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|error" gpurun_out/sanitize2_mem.log | head -12
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests -m gpu -q --tb=line -p no:cacheprovider -x -k "merge_block or ffn_fused or layernorm_loader" > gpurun_out/sanitize2_race.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/sanitize2_race.log | head -12
