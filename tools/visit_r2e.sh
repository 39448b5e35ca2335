#!/bin/bash
# GPU visit 5 of round 2: TMA-store epilogues (parity + timing), attention token variants + role timeline, bench.
TAG=${1:-r2e}
O=gpurun_out
mkdir -p $O
L=$O/exp_$TAG.log
: > $L
run() { name=$1; shift; echo "=== $name: $*" >> $L; timeout -s KILL ${T:-150} "$@" >> $L 2>&1; rc=$?; echo "=== $name rc=$rc" >> $L; return $rc; }
PT="python -m pytest -q -p no:cacheprovider --timeout 100 tests/test_kernels_gpu.py"
T=200 run t_gemm_tma $PT -k "gemm or conv3x3"; TMA=$?
VDB_EPI_TMA=0 T=200 run t_gemm_old $PT -k "gemm or conv3x3"
run mb_gemm_tma python tools/microbench.py gemm,conv $O/mb_gemm_tma_$TAG.json
VDB_EPI_TMA=0 run mb_gemm_old python tools/microbench.py gemm,conv $O/mb_gemm_old_$TAG.json
OKV=""
for v in 12 13 22; do
  VDB_ATT_FA=$v T=120 run t_fa$v $PT -k "attention" && OKV="$OKV $v"
done
for v in 11 $OKV; do
  VDB_ATT_FA=$v run mb_fa$v python tools/microbench.py attention $O/mb_fa${v}_$TAG.json
done
if [ -f tools/bin/libvdb200_tl.so ]; then
  export VDB200_LIB=$PWD/tools/bin/libvdb200_tl.so
  for v in 11 12 10; do VDB_ATT_FA=$v T=60 run tl_fa$v python tools/attention_fa_timeline.py; done
  unset VDB200_LIB
fi
cp $L $O/exp_$TAG.partial.log
if [ "$TMA" = "0" ]; then E=""; else E="VDB_EPI_TMA=0"; fi
export VDB_UPFOLD=1
T=500 run t_parity env $E python -m pytest -q -p no:cacheprovider --timeout 300 tests/test_parity_gpu.py -k "not benchmark_shape"
T=300 run bench_c2 env $E python bench.py --no-cpu-baseline
T=300 run bench_c2_oldepi env VDB_EPI_TMA=0 python bench.py --no-cpu-baseline
T=400 run step_breakdown env $E python tools/step_breakdown.py 10
grep -E "^===|passed|failed|\"value\"" $L | cut -c1-300
python - "$TAG" <<'PY'
import json, glob, sys
for f in sorted(glob.glob("gpurun_out/mb_*_%s.json" % sys.argv[1])):
    for r in json.load(open(f))["results"]:
        print(f, r["name"], r.get("graph_us"), r.get("graph_tflops"), r.get("graph_gbs"))
PY
