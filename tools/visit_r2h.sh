#!/bin/bash
# GPU visit 8: column-split two-tile attention (fa2), tile-width model for igemm, bench.
TAG=${1:-r2h}
O=gpurun_out
mkdir -p $O
L=$O/exp_$TAG.log
: > $L
run() { name=$1; shift; echo "=== $name: $*" >> $L; timeout -s KILL ${T:-150} "$@" >> $L 2>&1; rc=$?; echo "=== $name rc=$rc" >> $L; return $rc; }
PT="python -m pytest -q -p no:cacheprovider --timeout 100 tests/test_kernels_gpu.py"
T=300 run t_kernels $PT; KOK=$?
OKV=""
for v in 11 1 31 20; do
  VDB_ATT_FA2=$v T=120 run t_fa2_$v $PT -k "attention" && OKV="$OKV $v"
done
run mb_all python tools/microbench.py gemm,conv,attention $O/mb_all_$TAG.json
for v in $OKV; do
  VDB_ATT_FA2=$v run mb_fa2_$v python tools/microbench.py attention $O/mb_fa2_${v}_$TAG.json
done
VDB_BN_MODEL=0 run mb_nomodel python tools/microbench.py gemm,conv $O/mb_nomodel_$TAG.json
T=150 run ncu_fa2 ncu --set full --clock-control none --import-source on -k regex:attention_fa2_kernel --launch-skip 3 --launch-count 1 \
    -f -o $O/att_fa2_$TAG python tools/microbench.py attention $O/mb_ncu_fa.json
cp $L $O/exp_$TAG.partial.log
export VDB_UPFOLD=1
if [ "$KOK" = "0" ]; then A=""; else A="VDB_ATT_FA2=0"; fi
T=500 run t_parity env $A python -m pytest -q -p no:cacheprovider --timeout 300 tests/test_parity_gpu.py -k "not benchmark_shape"
T=300 run bench_c2 env $A python bench.py --no-cpu-baseline
T=400 run step_breakdown env $A python tools/step_breakdown.py 10
grep -E "^===|passed|failed|\"value\"" $L | cut -c1-300
python - "$TAG" <<'PY'
import json, glob, sys
for f in sorted(glob.glob("gpurun_out/mb_*_%s.json" % sys.argv[1])):
    for r in json.load(open(f))["results"]:
        print(f, r["name"], r.get("graph_us"), r.get("graph_tflops"), r.get("graph_gbs"))
PY
