#!/bin/bash
# GPU visit: skinny CUDA-core GEMM for small operands (0-D diffuser): kernel tests, text-latent parity, full-size timing, breakdown.
TAG=${1:-r2x}
O=gpurun_out
mkdir -p $O
L=$O/exp_$TAG.log
: > $L
run() { name=$1; shift; echo "=== $name: $*" >> $L; timeout -s KILL ${T:-60} "$@" >> $L 2>&1; rc=$?; echo "=== $name rc=$rc" >> $L; return $rc; }
T=120 run t_kernels python -m pytest -q -p no:cacheprovider --timeout 60 tests/test_kernels_gpu.py -k "skinny or gemm or attention"
T=200 run t_text python -m pytest -q -p no:cacheprovider -s --timeout 150 tests/test_parity_gpu.py -k "text_latent"
T=200 run text_bench python tools/text_flow_bench.py
VDB_SKINNY=0 T=200 run text_bench_noskinny python tools/text_flow_bench.py
T=240 run text_breakdown python tools/step_breakdown.py 10 --text
grep -E "^===|passed|failed|parity\]|workload|one DDIM|sum of" $L | cut -c1-420
grep -A14 "^by family:" $L | head -16
