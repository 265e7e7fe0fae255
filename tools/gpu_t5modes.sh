#!/bin/bash
for m in 0 1 2 3 4 5; do echo "mode $m"; SEB200_T5_MODE=$m timeout 100 python tools/attn_tc_check.py big 2>&1 | grep -E "variant 3" | grep -E "B=64|B=8 T=641 Fh=101 *variant 3:.*time" ; done
