#!/usr/bin/env bash
set -u
TAG="${1:-r01h}"
O=gpurun_out
mkdir -p $O
( time timeout 200 python bench.py --no-cpu ) > $O/${TAG}_bench_c3_devsummary.log 2>&1
( time B200_BENCH_VERBOSE=1 timeout 330 python bench.py --workload c4 --num-warmup 150 --steps 30 --no-cpu ) > $O/${TAG}_bench_c4.log 2>&1
ls -la $O | tail -4
