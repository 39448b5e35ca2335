#!/bin/bash
# GPU visit 19: CTA pairs (cta_group::2) again, now with the warp-uniform MMA issue and the TMA-store epilogue.
TAG=${1:-r2t}
O=gpurun_out
mkdir -p $O
L=$O/exp_$TAG.log
: > $L
run() { name=$1; shift; echo "=== $name: $*" >> $L; timeout -s KILL ${T:-60} "$@" >> $L 2>&1; rc=$?; echo "=== $name rc=$rc" >> $L; return $rc; }
VDB_PAIR=1 T=90 run t_pair python -m pytest -q -p no:cacheprovider --timeout 60 tests/test_kernels_gpu.py -k "gemm or conv3x3"
T=90 run mb_single python tools/microbench.py conv,gemm $O/mb_single_$TAG.json
VDB_PAIR=1 T=90 run mb_pair python tools/microbench.py conv,gemm $O/mb_pair_$TAG.json
VDB_PAIR=1 T=200 run bench_pair python bench.py --no-cpu-baseline
T=200 run bench_single python bench.py --no-cpu-baseline
grep -E "^===|passed|failed|\"value\"" $L | cut -c1-200
python - "$TAG" <<'PY'
import json, sys
a = {r["name"]: r for r in json.load(open("gpurun_out/mb_single_%s.json" % sys.argv[1]))["results"]}
b = {r["name"]: r for r in json.load(open("gpurun_out/mb_pair_%s.json" % sys.argv[1]))["results"]}
for k in a:
    if k in b:
        print(f"{k:34s} single {a[k].get('graph_us')} us  pair {b[k].get('graph_us')} us")
PY
