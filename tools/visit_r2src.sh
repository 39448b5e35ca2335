#!/bin/bash
# source-level ncu capture (stall samples per line) of igemm launches 5..8 of a DDIM step: the first ResBlock's convs and the first
# SpatialTransformer's LayerNorm-producer / consumer GEMMs
TAG=${1:-r2src}
O=gpurun_out
mkdir -p $O
L=$O/exp_$TAG.log
: > $L
run() { name=$1; shift; echo "=== $name: $*" >> $L; timeout -s KILL ${T:-60} "$@" >> $L 2>&1; rc=$?; echo "=== $name rc=$rc" >> $L; return $rc; }
T=240 run ncu_src ncu --set full --import-source on --clock-control none --profile-from-start off -k "regex:igemm_kernel" --launch-skip 4 --launch-count 6 -f -o $O/r02_src_igemm_$TAG python tools/profile_step.py
grep -E "^===" $L; ls -la $O | grep $TAG
