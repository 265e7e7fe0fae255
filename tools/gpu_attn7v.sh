#!/bin/bash
# experiment builds of the hybrid attention kernel (variant 4), time axis n = 641
export ATTN_BENCH_SHORT=1
echo "== default build"; timeout 300 python tools/attn_bench.py 3 4 2>&1 | grep "variant" | grep -v check
for s in "$@"; do echo "== $s"; SEB200_LIB_SUFFIX=$s timeout 300 python tools/attn_bench.py 4 2>&1 | grep "variant\|rror" | grep -v check | head -8; done
