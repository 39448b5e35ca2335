#!/bin/bash
# GPU visit 14: text-latent parity + timing with the FCBlock emb projections batched into one GEMM; (BN, split-K) sweep.
TAG=${1:-r2n}
O=gpurun_out
mkdir -p $O
L=$O/exp_$TAG.log
: > $L
run() { name=$1; shift; echo "=== $name: $*" >> $L; timeout -s KILL ${T:-60} "$@" >> $L 2>&1; rc=$?; echo "=== $name rc=$rc" >> $L; return $rc; }
T=200 run t_text python -m pytest -q -p no:cacheprovider -s --timeout 150 tests/test_parity_gpu.py -k "text_latent"
T=200 run text_bench python tools/text_flow_bench.py
T=300 run splitk_sweep python tools/splitk_sweep.py
grep -E "^===|passed|failed|parity|workload|conv3x3|gemm|linear_small" $L | cut -c1-400
