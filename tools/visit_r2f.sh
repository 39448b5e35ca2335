#!/bin/bash
# GPU visit 6 of round 2: warp-uniform MMA issue (igemm + attention), shared-memory-staged GroupNorm for the 64x64 level, PDL re-test.
TAG=${1:-r2f}
O=gpurun_out
mkdir -p $O
L=$O/exp_$TAG.log
: > $L
run() { name=$1; shift; echo "=== $name: $*" >> $L; timeout -s KILL ${T:-150} "$@" >> $L 2>&1; rc=$?; echo "=== $name rc=$rc" >> $L; return $rc; }
PT="python -m pytest -q -p no:cacheprovider --timeout 100 tests/test_kernels_gpu.py"
T=300 run t_kernels $PT; KOK=$?
VDB_ATT_FA=0 T=120 run t_att_old $PT -k "attention"
run mb_all python tools/microbench.py gemm,conv,attention,groupnorm $O/mb_all_$TAG.json
for v in 1 21 31 10; do
  VDB_ATT_FA=$v run mb_fa$v python tools/microbench.py attention $O/mb_fa${v}_$TAG.json
done
VDB_GN_BIG=0 run mb_gn_nobig python tools/microbench.py groupnorm $O/mb_gn_nobig_$TAG.json
if [ -f tools/bin/libvdb200_tl.so ]; then
  export VDB200_LIB=$PWD/tools/bin/libvdb200_tl.so
  for v in 11 21; do VDB_ATT_FA=$v T=60 run tl_fa$v python tools/attention_fa_timeline.py; done
  unset VDB200_LIB
fi
VDB_ATT_FA=11 T=150 run ncu_fa11 ncu --set full --clock-control none --import-source on -k regex:attention_fa_kernel --launch-skip 3 --launch-count 1 \
    -f -o $O/att_fa11_$TAG python tools/microbench.py attention $O/mb_ncu_fa.json
cp $L $O/exp_$TAG.partial.log
export VDB_UPFOLD=1
T=500 run t_parity python -m pytest -q -p no:cacheprovider --timeout 300 tests/test_parity_gpu.py -k "not benchmark_shape"
T=300 run bench_c2 python bench.py --no-cpu-baseline
T=300 run bench_c2_pdl env VDB_PDL=1 python bench.py --no-cpu-baseline
T=400 run step_breakdown python tools/step_breakdown.py 10
grep -E "^===|passed|failed|\"value\"" $L | cut -c1-300
python - "$TAG" <<'PY'
import json, glob, sys
for f in sorted(glob.glob("gpurun_out/mb_*_%s.json" % sys.argv[1])):
    for r in json.load(open(f))["results"]:
        print(f, r["name"], r.get("graph_us"), r.get("graph_tflops"), r.get("graph_gbs"))
PY
