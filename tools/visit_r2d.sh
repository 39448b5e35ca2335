#!/bin/bash
# GPU visit 4 of round 2: attention after the MMA issue-order fix, new ABI/CLIP tests, bench c2 (full line) / c3 / c4.
TAG=${1:-r2d}
O=gpurun_out
mkdir -p $O
L=$O/exp_$TAG.log
: > $L
run() { name=$1; shift; echo "=== $name: $*" >> $L; timeout -s KILL ${T:-150} "$@" >> $L 2>&1; rc=$?; echo "=== $name rc=$rc" >> $L; return $rc; }
PT="python -m pytest -q -p no:cacheprovider --timeout 100 tests/test_kernels_gpu.py"
OKV=""
for v in 11 1 21 111 101; do
  VDB_ATT_FA=$v T=120 run t_fa$v $PT -k "attention" && OKV="$OKV $v"
done
for v in $OKV 10; do
  VDB_ATT_FA=$v run mb_fa$v python tools/microbench.py attention $O/mb_fa${v}_$TAG.json
done
T=120 run t_misc $PT -k "repack or layernorm or groupnorm"
T=300 run t_clip python -m pytest -q -p no:cacheprovider -s --timeout 250 tests/test_clip.py -m gpu
case " $OKV " in *" 1 "*)
  VDB_ATT_FA=1 T=150 run ncu_fa1 ncu --set full --clock-control none --import-source on -k regex:attention_fa_kernel --launch-skip 3 --launch-count 1 \
    -f -o $O/att_fa1_$TAG python tools/microbench.py attention $O/mb_ncu_fa.json ;;
esac
cp $L $O/exp_$TAG.partial.log
export VDB_UPFOLD=1
T=400 run bench_c2 python bench.py
T=240 run bench_c3 python bench.py --config c3 --no-cpu-baseline
T=240 run bench_c4 python bench.py --config c4 --no-cpu-baseline
T=400 run step_breakdown python tools/step_breakdown.py 10
grep -E "^===|passed|failed|\"value\"" $L | cut -c1-400
python - "$TAG" <<'PY'
import json, glob, sys
for f in sorted(glob.glob("gpurun_out/mb_*_%s.json" % sys.argv[1])):
    for r in json.load(open(f))["results"]:
        print(f, r["name"], r.get("graph_us"), r.get("graph_tflops"), r.get("graph_gbs"))
PY
