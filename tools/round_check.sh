#!/bin/bash
# One GPU-box visit that produces everything a round needs: the GPU parity suite, the bench line, the ncu launch
# list of one DDIM step + VAE decode, and one `--set full` capture of the first ResBlock/SpatialTransformer kernels.
# Every leg has its own hard timeout; logs and reports land in gpurun_out/ (copy the summaries into profiles/).
TAG=${1:-v5}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$TAG.txt 2>&1
timeout -s KILL ${T_TEST:-420} python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 150 > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_$TAG.log
timeout -s KILL ${T_BENCH:-300} python bench.py > gpurun_out/bench_$TAG.log 2>&1
echo "bench exit $?" >> gpurun_out/bench_$TAG.log
if [ -z "$SKIP_NCU" ]; then
timeout -s KILL 200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/launches_$TAG.csv python tools/profile_step.py --decode > gpurun_out/ncu_launches_$TAG.log 2>&1
timeout -s KILL 200 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k "regex:attention_kernel|igemm_kernel" --launch-count 22 -f -o gpurun_out/r01_full_$TAG \
  python tools/profile_step.py > gpurun_out/ncu_full_$TAG.log 2>&1
fi
tail -3 gpurun_out/pytest_gpu_$TAG.log; tail -c 600 gpurun_out/bench_$TAG.log
