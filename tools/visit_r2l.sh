#!/bin/bash
# GPU visit 12 (2 GPUs, charged 2x): the 2-rank NCCL parity test and the bench's own torchrun launch at N=2 (both arms).
TAG=${1:-r2l}
O=gpurun_out
mkdir -p $O
L=$O/exp_$TAG.log
: > $L
run() { name=$1; shift; echo "=== $name: $*" >> $L; timeout -s KILL ${T:-60} "$@" >> $L 2>&1; rc=$?; echo "=== $name rc=$rc" >> $L; return $rc; }
nvidia-smi -L >> $L 2>&1
T=240 run t_dist python -m pytest -q -p no:cacheprovider --timeout 220 tests/test_dist_gpu.py
T=240 run bench_n2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline
T=120 run bench_n1 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline
T=200 run bench_ref_n2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --impl reference --gpus 2 --steps 1 --warmup 0
grep -E "^===|passed|failed|skipped|\"metric\"|\"impl\"" $L | cut -c1-900
