#!/bin/bash
# GPU visit 20: two-tile attention with the ones row: MUFU token off (mode 30), 4 K / V^T stages.
TAG=${1:-r2u}
O=gpurun_out
mkdir -p $O
L=$O/exp_$TAG.log
: > $L
run() { name=$1; shift; echo "=== $name: $*" >> $L; timeout -s KILL ${T:-60} "$@" >> $L 2>&1; rc=$?; echo "=== $name rc=$rc" >> $L; return $rc; }
T=60 run mb_default python tools/microbench.py attention $O/mb_att_default_$TAG.json
VDB_ATT_FA=30 T=45 run t_att_30 python -m pytest -q -p no:cacheprovider --timeout 40 tests/test_kernels_gpu.py -k attention
VDB_ATT_FA=30 T=60 run mb_notoken python tools/microbench.py attention $O/mb_att_notoken_$TAG.json
VDB_ATT_STAGES=4 T=45 run t_att_s4 python -m pytest -q -p no:cacheprovider --timeout 40 tests/test_kernels_gpu.py -k attention
VDB_ATT_STAGES=4 T=60 run mb_stages4 python tools/microbench.py attention $O/mb_att_stages4_$TAG.json
grep -E "^===|passed|failed" $L | cut -c1-200
python - "$TAG" <<'PY'
import json, glob, sys
for f in sorted(glob.glob("gpurun_out/mb_att_*_%s.json" % sys.argv[1])):
    for r in json.load(open(f))["results"][:1]:
        print(f, r["name"], r.get("graph_us"), r.get("graph_tflops"))
PY
