#!/bin/bash
# GPU visit 2 of round 2: parity + timing of the two-tile attention kernel (attention_fa_kernel) in its variants, the
# row-group LayerNorm, the repaired step_breakdown, and a bench with the validated switches.
TAG=${1:-r2b}
O=gpurun_out
mkdir -p $O
L=$O/exp_$TAG.log
: > $L
run() { name=$1; shift; echo "=== $name: $*" >> $L; timeout -s KILL ${T:-150} "$@" >> $L 2>&1; rc=$?; echo "=== $name rc=$rc" >> $L; return $rc; }
PT="python -m pytest -q -p no:cacheprovider --timeout 100 tests/test_kernels_gpu.py"
T=120 run t_ln $PT -k "layernorm"
OKV=""
for v in 21 1 20 31 41 11; do
  VDB_ATT_FA=$v T=120 run t_fa$v $PT -k "attention" && OKV="$OKV $v"
done
VDB_ATT_FA=0 run mb_fa0 python tools/microbench.py attention $O/mb_fa0_$TAG.json
for v in $OKV; do
  VDB_ATT_FA=$v run mb_fa$v python tools/microbench.py attention $O/mb_fa${v}_$TAG.json
done
run mb_ln python tools/microbench.py layernorm $O/mb_ln_$TAG.json
for v in 21; do
  case " $OKV " in *" $v "*)
  VDB_ATT_FA=$v T=150 run ncu_fa$v ncu --set full --clock-control none --import-source on -k regex:attention_fa_kernel --launch-skip 3 --launch-count 1 \
    -f -o $O/att_fa${v}_$TAG python tools/microbench.py attention $O/mb_ncu_fa.json ;;
  esac
done
T=400 run step_breakdown python tools/step_breakdown.py 10
cp $L $O/exp_$TAG.partial.log
T=240 run bench_default python bench.py --no-cpu-baseline
T=240 run bench_upfold env VDB_UPFOLD=1 python bench.py --no-cpu-baseline
T=600 run t_parity python -m pytest -q -p no:cacheprovider --timeout 300 tests/test_parity_gpu.py -k "not benchmark_shape"
grep -E "^===|passed|failed|\"value\"" $L | cut -c1-260
python - "$TAG" <<'PY'
import json, glob, sys
for f in sorted(glob.glob("gpurun_out/mb_*_%s.json" % sys.argv[1])):
    for r in json.load(open(f))["results"]:
        print(f, r["name"], r.get("graph_us"), r.get("graph_tflops"), r.get("graph_gbs"))
PY
