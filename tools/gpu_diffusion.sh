#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -k "merge or diffusion or predict_tsc or reverse_steps or gemm_rows_resid or layernorm_loader" > gpurun_out/pytest_diff.log 2>&1
echo "pytest rc=$?"; tail -n 25 gpurun_out/pytest_diff.log
timeout 600 python tools/diffusion_step.py > gpurun_out/diffusion_step.json 2> gpurun_out/diffusion_step.err
echo "step rc=$?"; tail -n 5 gpurun_out/diffusion_step.err; cat gpurun_out/diffusion_step.json
