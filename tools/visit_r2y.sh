#!/bin/bash
# Last visit of round 2 (final build): GPU suite, smoke, bench C2 (with CPU baseline) / C3 / C4, attention microbench, and a
# source-level ncu capture of the two-tile attention kernel and the GEGLU LayerNorm-consumer GEMM (stall samples per line).
TAG=${1:-r2y}
O=gpurun_out
mkdir -p $O
L=$O/exp_$TAG.log
: > $L
run() { name=$1; shift; echo "=== $name: $*" >> $L; timeout -s KILL ${T:-60} "$@" >> $L 2>&1; rc=$?; echo "=== $name rc=$rc" >> $L; return $rc; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $L 2>&1
T=500 run t_gpu python -m pytest -q -p no:cacheprovider --timeout 400 tests -m gpu
T=200 run smoke python -c "import __graft_entry__ as g; g.smoke()"
cp $L $O/exp_$TAG.partial.log
T=400 run bench_c2 python bench.py
T=60 run mb_att python tools/microbench.py attention $O/mb_att_$TAG.json
T=200 run bench_c3 python bench.py --config c3 --no-cpu-baseline
T=200 run bench_c4 python bench.py --config c4 --no-cpu-baseline
cp $L $O/exp_$TAG.partial.log
T=240 run ncu_src ncu --set full --import-source on --clock-control none --profile-from-start off -k "regex:attention_fa_kernel|igemm_kernel<256, 4, 1, 8, 6>" --launch-count 2 -f -o $O/r02_src_$TAG python tools/profile_step.py
T=300 run t_variants env VDB_TEST_VARIANTS=1 python -m pytest -q -p no:cacheprovider --timeout 280 tests/test_variants_gpu.py
grep -E "^===|passed|failed|\"value\"|smoke\]" $L | cut -c1-260
python - "$TAG" <<'PY'
import json, sys
for r in json.load(open("gpurun_out/mb_att_%s.json" % sys.argv[1]))["results"]:
    print(r["name"], r.get("graph_us"), r.get("graph_tflops"))
PY
ls -la $O | grep $TAG
