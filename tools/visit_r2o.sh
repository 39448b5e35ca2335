#!/bin/bash
# GPU visit 15: attention row sums through the ones row of V^T (ONES) — kernel tests, microbench A/B (incl. poly 2 / 3), bench.
TAG=${1:-r2o}
O=gpurun_out
mkdir -p $O
L=$O/exp_$TAG.log
: > $L
run() { name=$1; shift; echo "=== $name: $*" >> $L; timeout -s KILL ${T:-60} "$@" >> $L 2>&1; rc=$?; echo "=== $name rc=$rc" >> $L; return $rc; }
T=60 run t_att python -m pytest -q -p no:cacheprovider --timeout 45 tests/test_kernels_gpu.py -k attention
T=60 run mb_ones python tools/microbench.py attention $O/mb_att_ones_$TAG.json
VDB_ATT_ONES=0 T=60 run mb_noones python tools/microbench.py attention $O/mb_att_noones_$TAG.json
VDB_ATT_FA=21 T=60 run mb_ones_p2 python tools/microbench.py attention $O/mb_att_ones21_$TAG.json
VDB_ATT_FA=31 T=60 run mb_ones_p3 python tools/microbench.py attention $O/mb_att_ones31_$TAG.json
T=200 run bench_c2 python bench.py --no-cpu-baseline
grep -E "^===|passed|failed|\"value\"" $L | cut -c1-300
python - "$TAG" <<'PY'
import json, glob, sys
for f in sorted(glob.glob("gpurun_out/mb_*_%s.json" % sys.argv[1])):
    for r in json.load(open(f))["results"]:
        print(f, r["name"], r.get("graph_us"), r.get("graph_tflops"), r.get("graph_gbs"))
PY
