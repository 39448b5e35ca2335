#!/bin/bash
# GPU visit 16+: LayerNorm folded into the GEMM epilogues (vdb_gemm_ln_bf16) — kernel tests, path parity, bench A/B, breakdown.
TAG=${1:-r2p}
O=gpurun_out
mkdir -p $O
L=$O/exp_$TAG.log
: > $L
run() { name=$1; shift; echo "=== $name: $*" >> $L; timeout -s KILL ${T:-60} "$@" >> $L 2>&1; rc=$?; echo "=== $name rc=$rc" >> $L; return $rc; }
T=90 run t_gemm_ln python -m pytest -q -p no:cacheprovider --timeout 45 tests/test_kernels_gpu.py -k "gemm_ln or attention"
T=240 run t_parity python -m pytest -q -p no:cacheprovider -s --timeout 200 tests/test_parity_gpu.py -k "layernorm_fold or golden or oracle_fresh or graph_equals"
cp $L $O/exp_$TAG.partial.log
T=200 run bench_c2 python bench.py --no-cpu-baseline
VDB_LN_FOLD=0 T=200 run bench_c2_nofold python bench.py --no-cpu-baseline
T=200 run bench_c2_p41 env VDB_ATT_FA=41 python bench.py --no-cpu-baseline
T=300 run step_breakdown python tools/step_breakdown.py 10
grep -E "^===|passed|failed|\"value\"|parity\]" $L | cut -c1-260
