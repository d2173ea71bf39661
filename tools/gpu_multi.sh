#!/usr/bin/env bash
# Two-GPU box session (gpurun --gpus 2): multi-GPU parity tests, chain-sharded c3 (weak scaling, no collective) and
# row-sharded c5 (one ncclAllReduce per leapfrog) bench lines.
set -u
TAG="${1:-r01f}"
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv > $O/${TAG}_smi.txt 2>&1
( time timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -q ) > $O/${TAG}_pytest_multi.log 2>&1
echo "pytest rc=$?" >> $O/${TAG}_pytest_multi.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 \
   --num-warmup 300 --steps 100 --no-cpu > $O/${TAG}_bench_c3_n2.log 2>&1
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 \
   --workload c5 --scale 0.25 --num-warmup 100 --steps 16 --no-cpu > $O/${TAG}_bench_c5q_n2.log 2>&1
ls -la $O | tail -6
