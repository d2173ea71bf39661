#!/usr/bin/env bash
set -u
TAG="${1:-r01i}"
O=gpurun_out
mkdir -p $O
( time timeout 120 python bench.py --workload c2 --steps 200 ) > $O/${TAG}_bench_c2.log 2>&1
( time timeout 60 python bench.py --workload c1 --steps 1000 --no-cpu ) > $O/${TAG}_bench_c1.log 2>&1
ls -la $O | tail -3
