#!/usr/bin/env bash
# Round 2, second hardware session: the product-form logistic link (IRT + GLM DMMA v2) -- parity, then c3 / c4 / c5 numbers.
set -u
TAG="${1:-r02b}"
O=gpurun_out
mkdir -p $O
( time timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_posterior.py tests/test_gpu_summary.py -m gpu -q -x -p no:cacheprovider ) > $O/${TAG}_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/${TAG}_pytest_gpu.log
timeout 240 python bench.py --num-warmup 150 --steps 60 --no-cpu > $O/${TAG}_c3.log 2>&1
for V in 1 2; do
  B200_BENCH_VERBOSE=1 B200NUTS_IRT_VARIANT=$V timeout 240 python bench.py --workload c4 --chains 1024 --num-warmup 100 --steps 25 --no-cpu > $O/${TAG}_c4_irt$V.log 2>&1
done
for C in 64 128 256; do
  timeout 300 python bench.py --workload c5 --chains $C --num-warmup 100 --steps 16 --no-cpu > $O/${TAG}_c5_chains$C.log 2>&1
done
ls -la $O | tail -12
