#!/bin/bash
# Closing visit: final build (split-K reduction with all partials in flight): GPU suite, smoke, bench C2, step breakdown.
TAG=${1:-r2final}
O=gpurun_out
mkdir -p $O
L=$O/exp_$TAG.log
: > $L
run() { name=$1; shift; echo "=== $name: $*" >> $L; timeout -s KILL ${T:-60} "$@" >> $L 2>&1; rc=$?; echo "=== $name rc=$rc" >> $L; return $rc; }
T=500 run t_gpu python -m pytest -q -p no:cacheprovider --timeout 400 tests -m gpu
T=200 run smoke python -c "import __graft_entry__ as g; g.smoke()"
cp $L $O/exp_$TAG.partial.log
T=400 run bench_c2 python bench.py
T=300 run step_breakdown python tools/step_breakdown.py 10
grep -E "^===|passed|failed|\"value\"|smoke\]|one DDIM|sum of|splitk" $L | cut -c1-260
grep -A8 "^by family:" $L | head -10
