#!/bin/bash
# GPU visit 10: cleaned-up kernels (fa default 11), PDL re-test, round-1 attention kernel with the fast MMA issue, full GPU suite.
TAG=${1:-r2j}
O=gpurun_out
mkdir -p $O
L=$O/exp_$TAG.log
: > $L
run() { name=$1; shift; echo "=== $name: $*" >> $L; timeout -s KILL ${T:-60} "$@" >> $L 2>&1; rc=$?; echo "=== $name rc=$rc" >> $L; return $rc; }
T=120 run t_kernels python -m pytest -q -p no:cacheprovider --timeout 60 tests/test_kernels_gpu.py
T=90 run mb_att python tools/microbench.py attention $O/mb_att_$TAG.json
VDB_ATT_FA=0 T=90 run mb_att_r1 python tools/microbench.py attention $O/mb_att_r1_$TAG.json
T=700 run t_gpu_suite python -m pytest -q -p no:cacheprovider --timeout 400 tests -m gpu --deselect tests/test_kernels_gpu.py
cp $L $O/exp_$TAG.partial.log
T=300 run bench_c2 python bench.py --no-cpu-baseline
T=300 run bench_c2_pdl env VDB_PDL=1 python bench.py --no-cpu-baseline
T=300 run bench_c2_pdl_b env VDB_PDL=1 python bench.py --no-cpu-baseline --steps 5
T=300 run step_breakdown python tools/step_breakdown.py 10
T=200 run smoke python -c "import __graft_entry__ as g; g.smoke()"
grep -E "^===|passed|failed|\"value\"|smoke" $L | cut -c1-300
python - "$TAG" <<'PY'
import json, glob, sys
for f in sorted(glob.glob("gpurun_out/mb_*_%s.json" % sys.argv[1])):
    for r in json.load(open(f))["results"]:
        print(f, r["name"], r.get("graph_us"), r.get("graph_tflops"), r.get("graph_gbs"))
PY
