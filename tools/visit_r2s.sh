#!/bin/bash
# GPU visit 17: folded-LayerNorm step breakdown + the fold A/B parity test.
TAG=${1:-r2s}
O=gpurun_out
mkdir -p $O
L=$O/exp_$TAG.log
: > $L
run() { name=$1; shift; echo "=== $name: $*" >> $L; timeout -s KILL ${T:-60} "$@" >> $L 2>&1; rc=$?; echo "=== $name rc=$rc" >> $L; return $rc; }
T=120 run t_parity python -m pytest -q -p no:cacheprovider -s --timeout 100 tests/test_parity_gpu.py -k "layernorm_fold"
T=300 run step_breakdown python tools/step_breakdown.py 10
grep -E "^===|passed|failed|parity\]|one DDIM|sum of" $L | cut -c1-260
