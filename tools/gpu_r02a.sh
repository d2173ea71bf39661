#!/usr/bin/env bash
# Round 2, first hardware session (rationed: ~15 box-minutes): the whole GPU suite on HEAD including the kernels written
# after round 1's last session, then the A/B runs that decide what becomes the default.
# Run as: gpurun --timeout 1500 -- bash tools/gpu_r02a.sh [tag]
set -u
TAG="${1:-r02a}"
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_gpu.txt 2>&1
nproc >> $O/${TAG}_gpu.txt
( time B200_TEST_EXPERIMENTAL=1 B200_TEST_OUT_OF_SCOPE=1 timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider ) > $O/${TAG}_pytest_gpu_all.log 2>&1
echo "pytest rc=$?" >> $O/${TAG}_pytest_gpu_all.log
for V in 1 3; do
  B200NUTS_GLM_VARIANT=$V timeout 240 python bench.py --num-warmup 150 --steps 60 --no-cpu > $O/${TAG}_c3_glm$V.log 2>&1
  B200NUTS_GLM_VARIANT=$V timeout 240 python bench.py --workload c5 --scale 0.25 --num-warmup 100 --steps 16 --no-cpu > $O/${TAG}_c5q_glm$V.log 2>&1
done
for T in 0 nf 1; do
  B200_BENCH_VERBOSE=1 B200NUTS_IRT_TAIL=$T timeout 240 python bench.py --workload c4 --chains 1024 --num-warmup 100 --steps 25 --no-cpu > $O/${TAG}_c4_tail_$T.log 2>&1
done
B200_BENCH_VERBOSE=1 B200NUTS_IRT_TAIL=1 timeout 300 python bench.py --workload c4 --no-cpu > $O/${TAG}_c4_full_tail.log 2>&1
ls -la $O | tail -15
