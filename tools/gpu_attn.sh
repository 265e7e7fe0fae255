#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -k "attention" 2>&1 | tail -8
timeout 600 python tools/attn_ab.py 2>&1 | tail -20
