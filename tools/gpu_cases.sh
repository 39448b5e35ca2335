#!/bin/bash
# Runs every parametrised case matching $1 in its own process (sticky CUDA errors stay isolated).
mkdir -p gpurun_out
ids=$(python -m pytest ${2:-tests/test_kernels_gpu.py} --collect-only -q -k "$1" -p no:cacheprovider 2>/dev/null | grep "::" )
for id in $ids; do
  echo "=== $id" | tee -a gpurun_out/cases.log
  CUDA_LAUNCH_BLOCKING=1 timeout -s KILL ${3:-120} python -m pytest "$id" -q -x -s -p no:cacheprovider 2>&1 | grep -E "^E  |passed|failed|Error|parity\]" | head -14 | tee -a gpurun_out/cases.log
done
