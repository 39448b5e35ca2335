#!/bin/bash
# GPU visit 7: MUFU issue probe, phase-split attention body.
TAG=${1:-r2g}
O=gpurun_out
mkdir -p $O
L=$O/exp_$TAG.log
: > $L
run() { name=$1; shift; echo "=== $name: $*" >> $L; timeout -s KILL ${T:-150} "$@" >> $L 2>&1; rc=$?; echo "=== $name rc=$rc" >> $L; return $rc; }
PT="python -m pytest -q -p no:cacheprovider --timeout 100 tests/test_kernels_gpu.py"
T=60 run mufu_mix tools/bin/mufu_mix_bench
OKV=""
for v in 21 20 11 31 1021; do
  VDB_ATT_FA=$v T=120 run t_fa$v $PT -k "attention" && OKV="$OKV $v"
done
for v in $OKV; do
  VDB_ATT_FA=$v run mb_fa$v python tools/microbench.py attention $O/mb_fa${v}_$TAG.json
done
if [ -f tools/bin/libvdb200_tl.so ]; then
  export VDB200_LIB=$PWD/tools/bin/libvdb200_tl.so
  for v in 21 20; do VDB_ATT_FA=$v T=60 run tl_fa$v python tools/attention_fa_timeline.py; done
  unset VDB200_LIB
fi
grep -E "^===|passed|failed|\"value\"|mix" $L | cut -c1-300
python - "$TAG" <<'PY'
import json, glob, sys
for f in sorted(glob.glob("gpurun_out/mb_*_%s.json" % sys.argv[1])):
    for r in json.load(open(f))["results"]:
        print(f, r["name"], r.get("graph_us"), r.get("graph_tflops"), r.get("graph_gbs"))
PY
