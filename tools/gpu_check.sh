#!/bin/bash
# Runs the GPU kernel tests group by group, each under its own hard timeout (a hung kernel must not
# take the whole box lease with it). Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for grp in "$@"; do
  echo "=== $grp" | tee -a gpurun_out/check.log
  timeout -s KILL 420 python -m pytest tests/test_kernels_gpu.py -q -k "$grp" -p no:cacheprovider 2>&1 | tail -40 | tee -a gpurun_out/check.log
done
