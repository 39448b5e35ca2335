#!/bin/bash
# GPU visit 21: attention with the tail mask split out of the steady-state tile body (630 instead of 1026 instructions per tile).
TAG=${1:-r2v}
O=gpurun_out
mkdir -p $O
L=$O/exp_$TAG.log
: > $L
run() { name=$1; shift; echo "=== $name: $*" >> $L; timeout -s KILL ${T:-60} "$@" >> $L 2>&1; rc=$?; echo "=== $name rc=$rc" >> $L; return $rc; }
T=60 run t_att python -m pytest -q -p no:cacheprovider --timeout 45 tests/test_kernels_gpu.py -k attention
T=60 run mb_p3 python tools/microbench.py attention $O/mb_att_p3_$TAG.json
VDB_ATT_FA=11 T=60 run mb_p1 python tools/microbench.py attention $O/mb_att_p1_$TAG.json
VDB_ATT_FA=21 T=60 run mb_p2 python tools/microbench.py attention $O/mb_att_p2_$TAG.json
VDB_ATT_FA=41 T=60 run mb_p4 python tools/microbench.py attention $O/mb_att_p4_$TAG.json
VDB_ATT_ONES=0 T=60 run mb_noones python tools/microbench.py attention $O/mb_att_noones_$TAG.json
T=120 run t_clip python -m pytest -q -p no:cacheprovider --timeout 100 tests/test_clip.py -m gpu
T=200 run bench_c2 python bench.py --no-cpu-baseline
grep -E "^===|passed|failed|\"value\"" $L | cut -c1-200
python - "$TAG" <<'PY'
import json, glob, sys
for f in sorted(glob.glob("gpurun_out/mb_att_*_%s.json" % sys.argv[1])):
    for r in json.load(open(f))["results"]:
        print(f, r["name"], r.get("graph_us"), r.get("graph_tflops"))
PY
