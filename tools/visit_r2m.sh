#!/bin/bash
# GPU visit 13: in-graph breakdown of the text-latent (i2t) DDIM step, warm per launch and cold per family.
TAG=${1:-r2m}
O=gpurun_out
mkdir -p $O
L=$O/exp_$TAG.log
: > $L
run() { name=$1; shift; echo "=== $name: $*" >> $L; timeout -s KILL ${T:-60} "$@" >> $L 2>&1; rc=$?; echo "=== $name rc=$rc" >> $L; return $rc; }
T=240 run text_breakdown python tools/step_breakdown.py 10 --text
tail -60 $L | cut -c1-200
