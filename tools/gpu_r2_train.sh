#!/bin/bash
# round 2: the training-step tests, one pytest process per test function (a faulting kernel must not mask the others)
mkdir -p gpurun_out
: > gpurun_out/r2_train.log
for t in $(python -m pytest tests/test_gpu_train.py --collect-only -q -m gpu 2>/dev/null | grep "::" | sed 's/\[.*//' | sort -u); do
  echo "=== $t" >> gpurun_out/r2_train.log
  timeout 600 python -m pytest "$t" -q --tb=short -p no:cacheprovider -s 2>&1 | grep -v "^$" | tail -${TAILN:-25} >> gpurun_out/r2_train.log
done
grep -E "^===|passed|failed|Error|error|assert|\[simt\]|\[tcgen05" gpurun_out/r2_train.log | cut -c1-300 | tail -150
