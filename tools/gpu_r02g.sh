#!/usr/bin/env bash
# Round 2, two-GPU session: multi-rank parity tests, then bench.py exactly as the driver launches it for N = 2 (c3 headline, chains
# sharded; attached: c4 8192 chains split over the GPUs, c5 rows split with one ncclAllReduce per leapfrog).
# Run as: gpurun --gpus 2 --timeout 1500 -- bash tools/gpu_r02g.sh
set -u
TAG="${1:-r02g}"
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv > $O/${TAG}_gpus.txt 2>&1
( time timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -p no:cacheprovider ) > $O/${TAG}_pytest_multi.log 2>&1
echo "pytest rc=$?" >> $O/${TAG}_pytest_multi.log
( time B200_BENCH_VERBOSE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 20 --warmup 5 ) > $O/${TAG}_bench_n2.json 2> $O/${TAG}_bench_n2.err
echo "bench rc=$?" >> $O/${TAG}_bench_n2.err
ls -la $O | tail -6
