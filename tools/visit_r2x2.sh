#!/bin/bash
# GPU visit: skinny GEMM after the staging fix: kernel tests, text parity, timing at thresholds 0 / 16 / 64.
TAG=${1:-r2x2}
O=gpurun_out
mkdir -p $O
L=$O/exp_$TAG.log
: > $L
run() { name=$1; shift; echo "=== $name: $*" >> $L; timeout -s KILL ${T:-60} "$@" >> $L 2>&1; rc=$?; echo "=== $name rc=$rc" >> $L; return $rc; }
T=120 run t_kernels python -m pytest -q -p no:cacheprovider --timeout 60 tests/test_kernels_gpu.py -k "skinny"
T=200 run t_text python -m pytest -q -p no:cacheprovider -s --timeout 150 tests/test_parity_gpu.py -k "text_latent"
VDB_SKINNY=16 T=200 run text_bench_16 python tools/text_flow_bench.py
VDB_SKINNY=64 T=200 run text_bench_64 python tools/text_flow_bench.py
VDB_SKINNY=0 T=200 run text_bench_0 python tools/text_flow_bench.py
VDB_SKINNY=16 T=240 run text_breakdown python tools/step_breakdown.py 10 --text
grep -E "^===|passed|failed|parity\]|workload|one DDIM|sum of|gemm_skinny S" $L | cut -c1-420
