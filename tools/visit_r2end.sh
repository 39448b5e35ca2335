#!/bin/bash
# after reverting the split-K reduction experiment: the kernel tests that cover it + one bench line
TAG=${1:-r2end}
O=gpurun_out
mkdir -p $O
L=$O/exp_$TAG.log
: > $L
run() { name=$1; shift; echo "=== $name: $*" >> $L; timeout -s KILL ${T:-60} "$@" >> $L 2>&1; rc=$?; echo "=== $name rc=$rc" >> $L; return $rc; }
T=120 run t_kernels python -m pytest -q -p no:cacheprovider --timeout 60 tests/test_kernels_gpu.py
T=200 run bench_c2 python bench.py --no-cpu-baseline
grep -E "^===|passed|failed|\"value\"" $L | cut -c1-220
