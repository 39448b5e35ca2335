#!/bin/bash
# the binary that ships: kernel tests + smoke
TAG=${1:-r2last}
O=gpurun_out
mkdir -p $O
L=$O/exp_$TAG.log
: > $L
run() { name=$1; shift; echo "=== $name: $*" >> $L; timeout -s KILL ${T:-60} "$@" >> $L 2>&1; rc=$?; echo "=== $name rc=$rc" >> $L; return $rc; }
T=120 run t_kernels python -m pytest -q -p no:cacheprovider --timeout 60 tests/test_kernels_gpu.py
T=150 run smoke python -c "import __graft_entry__ as g; g.smoke()"
grep -E "^===|passed|failed|smoke\]" $L | cut -c1-200
