#!/bin/bash
# Final visit of round 2: GPU suite (+ opt-in variants), smoke, bench (C2 with the CPU baseline, C3, C4), ncu launch list of one DDIM
# step + decode, one --set full capture of the step's first tensor-core / attention / GroupNorm launches.
TAG=${1:-r2z}
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
T=200 run ncu_launches ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_$TAG.csv python tools/profile_step.py --decode
T=300 run ncu_full ncu --set full --clock-control none --profile-from-start off -k "regex:igemm_kernel|attention_fa_kernel|attention_kernel|gn_bundle_kernel" --launch-count 36 -f -o $O/r02_full_$TAG python tools/profile_step.py
cp $L $O/exp_$TAG.partial.log
T=200 run bench_c3 python bench.py --config c3 --no-cpu-baseline
T=200 run bench_c4 python bench.py --config c4 --no-cpu-baseline
T=400 run t_variants env VDB_TEST_VARIANTS=1 python -m pytest -q -p no:cacheprovider --timeout 380 tests/test_variants_gpu.py
grep -E "^===|passed|failed|\"value\"|smoke\]" $L | cut -c1-260
ls -la $O | grep $TAG
