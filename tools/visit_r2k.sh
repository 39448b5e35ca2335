#!/bin/bash
# GPU visit 11: text-latent flows (parity + full-size timing).
TAG=${1:-r2k}
O=gpurun_out
mkdir -p $O
L=$O/exp_$TAG.log
: > $L
run() { name=$1; shift; echo "=== $name: $*" >> $L; timeout -s KILL ${T:-60} "$@" >> $L 2>&1; rc=$?; echo "=== $name rc=$rc" >> $L; return $rc; }
T=200 run t_text python -m pytest -q -p no:cacheprovider -s --timeout 150 tests/test_parity_gpu.py -k "text_latent"
T=300 run text_bench python tools/text_flow_bench.py
grep -E "^===|passed|failed|parity|workload|built" $L | cut -c1-600
