#!/usr/bin/env bash
set -u
TAG="${1:-q}"
O=gpurun_out
mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/${TAG}_pytest_gpu.log 2>&1
tail -6 $O/${TAG}_pytest_gpu.log
( time timeout 900 python bench.py --no-cpu ) > $O/${TAG}_bench_c3.log 2>&1
tail -4 $O/${TAG}_bench_c3.log | cut -c1-600
( time timeout 900 python bench.py --workload c4 --chains 1024 --num-warmup 100 --steps 20 --no-cpu ) > $O/${TAG}_bench_c4_small.log 2>&1
tail -4 $O/${TAG}_bench_c4_small.log | cut -c1-600
( time timeout 900 python bench.py --workload c5 --scale 0.1 --num-warmup 300 --steps 30 --no-cpu ) > $O/${TAG}_bench_c5_small.log 2>&1
tail -4 $O/${TAG}_bench_c5_small.log | cut -c1-600
