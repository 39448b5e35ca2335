#!/bin/bash
# the opt-in kernel variants on the final build
TAG=${1:-r2var}
O=gpurun_out
mkdir -p $O
L=$O/exp_$TAG.log
: > $L
run() { name=$1; shift; echo "=== $name: $*" >> $L; timeout -s KILL ${T:-60} "$@" >> $L 2>&1; rc=$?; echo "=== $name rc=$rc" >> $L; return $rc; }
T=400 run t_variants env VDB_TEST_VARIANTS=1 python -m pytest -q -p no:cacheprovider --timeout 380 tests/test_variants_gpu.py
grep -E "^===|passed|failed" $L | cut -c1-200
