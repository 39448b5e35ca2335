#!/bin/bash
# folded upsample convs storing straight into the interleaved result (conv modes 7..10): kernel tests, path parity, bench A/B
TAG=${1:-r2up}
O=gpurun_out
mkdir -p $O
L=$O/exp_$TAG.log
: > $L
run() { name=$1; shift; echo "=== $name: $*" >> $L; timeout -s KILL ${T:-60} "$@" >> $L 2>&1; rc=$?; echo "=== $name rc=$rc" >> $L; return $rc; }
T=120 run t_kernels python -m pytest -q -p no:cacheprovider --timeout 60 tests/test_kernels_gpu.py -k "upsample or conv3x3"
T=240 run t_parity python -m pytest -q -p no:cacheprovider --timeout 200 tests/test_parity_gpu.py -k "golden or graph_equals or benchmark_shape"
T=200 run bench_direct python bench.py --no-cpu-baseline
VDB_UPFOLD_DIRECT=0 T=200 run bench_pass python bench.py --no-cpu-baseline
grep -E "^===|passed|failed|\"value\"" $L | cut -c1-220
