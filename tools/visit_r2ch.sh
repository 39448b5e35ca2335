#!/bin/bash
# contiguous tile ranges per CTA (p.chunked) vs the grid-strided walk: kernel tests, GEMM / conv microbench A/B, bench A/B
TAG=${1:-r2ch}
O=gpurun_out
mkdir -p $O
L=$O/exp_$TAG.log
: > $L
run() { name=$1; shift; echo "=== $name: $*" >> $L; timeout -s KILL ${T:-60} "$@" >> $L 2>&1; rc=$?; echo "=== $name rc=$rc" >> $L; return $rc; }
T=120 run t_kernels python -m pytest -q -p no:cacheprovider --timeout 60 tests/test_kernels_gpu.py -k "gemm or conv3x3 or upsample"
T=90 run mb_chunked python tools/microbench.py conv,gemm $O/mb_chunked_$TAG.json
VDB_CHUNKED=0 T=90 run mb_strided python tools/microbench.py conv,gemm $O/mb_strided_$TAG.json
T=200 run bench_chunked python bench.py --no-cpu-baseline
VDB_CHUNKED=0 T=200 run bench_strided python bench.py --no-cpu-baseline
grep -E "^===|passed|failed|\"value\"" $L | cut -c1-200
python - "$TAG" <<'PY'
import json, sys
a = {r["name"]: r for r in json.load(open("gpurun_out/mb_strided_%s.json" % sys.argv[1]))["results"]}
b = {r["name"]: r for r in json.load(open("gpurun_out/mb_chunked_%s.json" % sys.argv[1]))["results"]}
for k in a:
    if k in b:
        print(f"{k:34s} strided {a[k].get('graph_us')} us  chunked {b[k].get('graph_us')} us")
PY
