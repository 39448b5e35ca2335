#!/bin/bash
# First GPU visit of round 2: validate + measure the variants written blind at the end of round 1
# (ping-pong attention VDB_ATT_PP=2|3, N-fast GEMM tile order VDB_NFAST).  Every leg has its own timeout; a hung
# variant costs at most that leg.
TAG=${1:-r2a}
O=gpurun_out
mkdir -p $O
run() { name=$1; shift; echo "=== $name: $*" >> $O/exp_$TAG.log; timeout -s KILL ${T:-150} "$@" >> $O/exp_$TAG.log 2>&1; rc=$?; echo "=== $name rc=$rc" >> $O/exp_$TAG.log; return $rc; }
PT="python -m pytest -q -p no:cacheprovider --timeout 100 tests/test_kernels_gpu.py"
VDB_TEST_PENDING=1 T=300 run t_pending python -m pytest -q -p no:cacheprovider --timeout 200 tests/test_pending_gpu.py
T=400 run t_benchshape python -m pytest -q -p no:cacheprovider -s --timeout 380 tests/test_parity_gpu.py -k "benchmark_shape"
VDB_ATT_PP=3 T=90 run t_pp3 $PT -k "attention"; PP3=$?
VDB_ATT_PP=2 T=90 run t_pp2 $PT -k "attention"; PP2=$?
VDB_NFAST=2 T=120 run t_nfast $PT -k "gemm or conv3x3"; NF=$?
VDB_TEST_VARIANTS=1 T=400 run t_upfold python -m pytest -q -p no:cacheprovider tests/test_variants_gpu.py -k "folded"; UF=$?
VDB_GN_CLUSTER=5 T=90 run t_gncl $PT -k "groupnorm"; GC=$?
VDB_GN_CLUSTER=7 T=90 run t_gncl_nokeep $PT -k "groupnorm"
VDB_LN_V2=1 T=90 run t_lnv2 $PT -k "layernorm"; LN=$?
run mb_default python tools/microbench.py attention,gemm,groupnorm,layernorm $O/mb_default_$TAG.json
[ "$LN" = "0" ] && VDB_LN_V2=1 run mb_lnv2 python tools/microbench.py layernorm $O/mb_lnv2_$TAG.json
[ "$GC" = "0" ] && VDB_GN_CLUSTER=1 run mb_gncl python tools/microbench.py groupnorm $O/mb_gncl_$TAG.json
[ "$GC" = "0" ] && VDB_GN_CLUSTER=3 run mb_gncl_nokeep python tools/microbench.py groupnorm $O/mb_gncl_nokeep_$TAG.json
[ "$PP3" = "0" ] && VDB_ATT_PP=3 run mb_pp3 python tools/microbench.py attention $O/mb_pp3_$TAG.json
[ "$PP2" = "0" ] && VDB_ATT_PP=2 run mb_pp2 python tools/microbench.py attention $O/mb_pp2_$TAG.json
[ "$NF" = "0" ] && VDB_NFAST=1 run mb_nfast python tools/microbench.py gemm $O/mb_nfast_$TAG.json
if [ "$PP3" = "0" ]; then
  VDB_ATT_PP=3 T=120 run ncu_pp3 ncu --set full --clock-control none --import-source on -k regex:attention_pp_kernel --launch-skip 3 --launch-count 1 \
    -f -o $O/att_pp3_$TAG python tools/microbench.py attention $O/mb_ncu_pp3.json
fi
T=180 run step_breakdown python tools/step_breakdown.py 10
cp $O/exp_$TAG.log $O/exp_$TAG.partial.log 2>/dev/null
# one bench with every variant that passed its parity leg
FLAGS=""
[ "$PP3" = "0" ] && FLAGS="$FLAGS VDB_ATT_PP=3"
[ "$NF" = "0" ] && FLAGS="$FLAGS VDB_NFAST=1"
[ "$GC" = "0" ] && FLAGS="$FLAGS VDB_GN_CLUSTER=1"
[ "$UF" = "0" ] && FLAGS="$FLAGS VDB_UPFOLD=1"
[ "$LN" = "0" ] && FLAGS="$FLAGS VDB_LN_V2=1"
echo "=== bench flags:$FLAGS" >> $O/exp_$TAG.log
T=240 run bench_all env $FLAGS python bench.py --no-cpu-baseline
if [ -f tools/bin/libvdb200_tl.so ]; then   # built HERE beforehand with tools/build_timeline_lib.sh (nvcc is on the box too, but slower)
  export VDB200_LIB=$PWD/tools/bin/libvdb200_tl.so
  T=40 run tl_gn_4096_320 python tools/gn_timeline.py 4096 320
  T=40 run tl_gn_1024_640 python tools/gn_timeline.py 1024 640
  T=40 run tl_gn_64_1280 python tools/gn_timeline.py 64 1280
  unset VDB200_LIB
fi
grep -E "^===|passed|failed|\"value\"" $O/exp_$TAG.log | cut -c1-260
python - "$TAG" <<'PY'
import json, glob, sys
for f in sorted(glob.glob("gpurun_out/mb_*_%s.json" % sys.argv[1])):
    for r in json.load(open(f))["results"]:
        if r["name"].startswith(("attention N4096 M4096", "gemm 32768x320x1280", "groupnorm", "layernorm")): print(f, r["name"], r.get("graph_us"))
PY
