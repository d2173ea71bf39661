#!/usr/bin/env bash
# ncu --set full captures of the secondary kernels (fused chain kernel, advance kernel, IRT gradient)
set -u
TAG="${1:-p}"
O=gpurun_out
mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_chain -s 1 -c 1 -o $O/${TAG}_fused \
   python bench.py --workload c2 --num-warmup 200 --steps 50 --warmup 3 --no-cpu > $O/${TAG}_ncu_fused_stdout.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:advance_kernel -s 300 -c 1 -o $O/${TAG}_advance \
   python bench.py --num-warmup 30 --steps 4 --warmup 3 --no-cpu > $O/${TAG}_ncu_advance_stdout.log 2>&1
( time timeout 900 python bench.py --workload c4 --chains 1024 --num-warmup 100 --steps 20 --no-cpu ) > $O/${TAG}_bench_c4_small.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:irt_grad -s 100 -c 1 -o $O/${TAG}_irt \
   python bench.py --workload c4 --chains 1024 --num-warmup 10 --steps 4 --warmup 3 --no-cpu > $O/${TAG}_ncu_irt_stdout.log 2>&1
ls -la $O | tail -8
