#!/bin/bash
# A/B visit to the GPU box for the opt-in kernel variants: correctness of each variant first (own process: the
# VDB_* switches are read once per process), then micro-benchmarks per variant, then one bench.py per candidate.
TAG=${1:-x1}
O=gpurun_out
mkdir -p $O
run() { name=$1; shift; echo "=== $name: $*" >> $O/exp_$TAG.log; timeout -s KILL ${T:-150} "$@" >> $O/exp_$TAG.log 2>&1; rc=$?; echo "=== $name rc=$rc" >> $O/exp_$TAG.log; return $rc; }
PT="python -m pytest -q -p no:cacheprovider --timeout 100 tests/test_kernels_gpu.py"
run t_default $PT -k "groupnorm or attention"
VDB_ATT_BKV=64 run t_bkv64 $PT -k "attention"; BKV_OK=$?
VDB_GN_REG=0 run t_gnreg0 $PT -k "groupnorm"
run mb_default python tools/microbench.py attention,groupnorm,layernorm $O/mb_default_$TAG.json
VDB_ATT_BKV=64 run mb_bkv64 python tools/microbench.py attention $O/mb_bkv64_$TAG.json
VDB_GN_REG=0 run mb_gnreg0 python tools/microbench.py groupnorm $O/mb_gnreg0_$TAG.json
VDB_GN_FUSED=0 run mb_gnfused0 python tools/microbench.py groupnorm $O/mb_gnfused0_$TAG.json
VDB_PAIR=1 run pair1 python tools/pair_check.py
VDB_PAIR=0 run pair0 python tools/pair_check.py
if [ "$BKV_OK" = "0" ]; then
  VDB_ATT_BKV=64 T=240 run bench_bkv64 python bench.py --no-cpu-baseline
else
  T=240 run bench_default python bench.py --no-cpu-baseline
fi
grep -E "^===|passed|failed|PAIR_CHECK|\"value\"" $O/exp_$TAG.log | cut -c1-300
