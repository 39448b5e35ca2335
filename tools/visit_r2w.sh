#!/bin/bash
# GPU visit: the new multi-image-context attention cases.
TAG=${1:-r2w}
O=gpurun_out
mkdir -p $O
L=$O/exp_$TAG.log
: > $L
run() { name=$1; shift; echo "=== $name: $*" >> $L; timeout -s KILL ${T:-60} "$@" >> $L 2>&1; rc=$?; echo "=== $name rc=$rc" >> $L; return $rc; }
T=90 run t_att python -m pytest -q -p no:cacheprovider --timeout 60 tests/test_kernels_gpu.py -k attention
grep -E "^===|passed|failed" $L | cut -c1-200
