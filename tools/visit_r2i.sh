#!/bin/bash
# GPU visit 9: column-split attention (fa2) after the setmaxnreg fix, GroupNorm one-way exchange.  Short legs: a hang costs 40 s.
TAG=${1:-r2i}
O=gpurun_out
mkdir -p $O
L=$O/exp_$TAG.log
: > $L
run() { name=$1; shift; echo "=== $name: $*" >> $L; timeout -s KILL ${T:-60} "$@" >> $L 2>&1; rc=$?; echo "=== $name rc=$rc" >> $L; return $rc; }
PT="python -m pytest -q -p no:cacheprovider --timeout 30 tests/test_kernels_gpu.py"
T=45 run t_fa2 $PT -k "attention"; FA2=$?
T=45 run t_gn $PT -k "groupnorm"; GN=$?
E=""
[ "$FA2" = "0" ] || E="$E VDB_ATT_FA2=0"
[ "$GN" = "0" ] || E="$E VDB_GN_BUNDLE=0"
if [ "$FA2" = "0" ]; then
  for v in 21 11 1 31 20; do
    VDB_ATT_FA2=$v T=45 run t_fa2_$v $PT -k "attention" && VDB_ATT_FA2=$v T=60 run mb_fa2_$v python tools/microbench.py attention $O/mb_fa2_${v}_$TAG.json
  done
  T=120 run ncu_fa2 ncu --set full --clock-control none --import-source on -k regex:attention_fa2_kernel --launch-skip 3 --launch-count 1 \
    -f -o $O/att_fa2_$TAG python tools/microbench.py attention $O/mb_ncu_fa.json
fi
T=90 run t_kernels env $E $PT
T=90 run mb_gn env $E python tools/microbench.py groupnorm $O/mb_gn_$TAG.json
cp $L $O/exp_$TAG.partial.log
T=300 run t_parity env $E python -m pytest -q -p no:cacheprovider --timeout 200 tests/test_parity_gpu.py -k "not benchmark_shape"
T=240 run bench_c2 env $E python bench.py --no-cpu-baseline
T=300 run step_breakdown env $E python tools/step_breakdown.py 10
grep -E "^===|passed|failed|\"value\"" $L | cut -c1-300
python - "$TAG" <<'PY'
import json, glob, sys
for f in sorted(glob.glob("gpurun_out/mb_*_%s.json" % sys.argv[1])):
    for r in json.load(open(f))["results"]:
        print(f, r["name"], r.get("graph_us"), r.get("graph_tflops"), r.get("graph_gbs"))
PY
