#!/bin/bash
# GPU visit 3 of round 2: P-in-tensor-memory attention variants, group-bundle GroupNorm, bench.
TAG=${1:-r2c}
O=gpurun_out
mkdir -p $O
L=$O/exp_$TAG.log
: > $L
run() { name=$1; shift; echo "=== $name: $*" >> $L; timeout -s KILL ${T:-150} "$@" >> $L 2>&1; rc=$?; echo "=== $name rc=$rc" >> $L; return $rc; }
PT="python -m pytest -q -p no:cacheprovider --timeout 100 tests/test_kernels_gpu.py"
T=150 run t_gn $PT -k "groupnorm"; GN=$?
OKV=""
for v in 11 1 21 31 10 20; do
  VDB_ATT_FA=$v T=120 run t_fa$v $PT -k "attention" && OKV="$OKV $v"
done
for v in 111 $OKV; do
  VDB_ATT_FA=$v run mb_fa$v python tools/microbench.py attention $O/mb_fa${v}_$TAG.json
done
run mb_gn python tools/microbench.py groupnorm $O/mb_gn_$TAG.json
VDB_GN_BUNDLE=0 run mb_gn_old python tools/microbench.py groupnorm $O/mb_gnold_$TAG.json
for v in 11; do
  case " $OKV " in *" $v "*)
  VDB_ATT_FA=$v T=150 run ncu_fa$v ncu --set full --clock-control none --import-source on -k regex:attention_fa_kernel --launch-skip 3 --launch-count 1 \
    -f -o $O/att_fa${v}_$TAG python tools/microbench.py attention $O/mb_ncu_fa.json ;;
  esac
done
[ "$GN" = "0" ] && T=150 run ncu_gn ncu --set full --clock-control none --import-source on -k regex:gn_bundle_kernel --launch-skip 3 --launch-count 2 \
    -f -o $O/gn_bundle_$TAG python tools/microbench.py groupnorm $O/mb_ncu_gn.json
cp $L $O/exp_$TAG.partial.log
if [ "$GN" = "0" ]; then B=""; else B="VDB_GN_BUNDLE=0"; fi
T=400 run t_parity env $B python -m pytest -q -p no:cacheprovider --timeout 300 tests/test_parity_gpu.py -k "not benchmark_shape and not c1_full"
T=240 run bench_default env $B VDB_UPFOLD=1 python bench.py --no-cpu-baseline
T=400 run step_breakdown env $B VDB_UPFOLD=1 python tools/step_breakdown.py 10
grep -E "^===|passed|failed|\"value\"" $L | cut -c1-260
python - "$TAG" <<'PY'
import json, glob, sys
for f in sorted(glob.glob("gpurun_out/mb_*_%s.json" % sys.argv[1])):
    for r in json.load(open(f))["results"]:
        print(f, r["name"], r.get("graph_us"), r.get("graph_tflops"), r.get("graph_gbs"))
PY
