#!/bin/bash
# attention variants: correctness, micro-benchmark, and one ncu --set full capture (N=4096, d=40 self-attention)
TAG=${1:-x3}
O=gpurun_out
mkdir -p $O
run() { name=$1; shift; echo "=== $name: $*" >> $O/exp_$TAG.log; timeout -s KILL ${T:-150} "$@" >> $O/exp_$TAG.log 2>&1; rc=$?; echo "=== $name rc=$rc" >> $O/exp_$TAG.log; return $rc; }
PT="python -m pytest -q -p no:cacheprovider --timeout 100 tests/test_kernels_gpu.py"
VDB_ATT_BKV=643 run t_bkv643 $PT -k "attention"
VDB_ATT_BKV=643 run mb_bkv643 python tools/microbench.py attention $O/mb_bkv643_$TAG.json
VDB_ATT_BKV=643 T=120 run ncu_643 ncu --set full --clock-control none --import-source on -k regex:attention_kernel --launch-skip 3 --launch-count 1 \
    -f -o $O/att_643_$TAG python tools/microbench.py attention $O/mb_ncu_643.json
grep -E "^===|passed|failed" $O/exp_$TAG.log | cut -c1-200
