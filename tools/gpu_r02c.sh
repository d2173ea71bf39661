#!/usr/bin/env bash
# Round 2, third hardware session: the IRT latency kernel (cluster of 8 CTAs per straggler chain) -- bit identity, then c4.
set -u
TAG="${1:-r02c}"
O=gpurun_out
mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "irt or lp_grad or lockstep" ) > $O/${TAG}_pytest_irt.log 2>&1
echo "pytest rc=$?" >> $O/${TAG}_pytest_irt.log
for S in 0 48; do
  B200_BENCH_VERBOSE=1 B200NUTS_IRT_SOLO=$S timeout 300 python bench.py --workload c4 --chains 1024 --num-warmup 100 --steps 25 --no-cpu --extras none > $O/${TAG}_c4_1024_solo$S.log 2>&1
done
B200_BENCH_VERBOSE=1 timeout 400 python bench.py --workload c4 --no-cpu --extras none > $O/${TAG}_c4_full.log 2>&1
B200_BENCH_VERBOSE=1 B200NUTS_IRT_SOLO=148 timeout 400 python bench.py --workload c4 --no-cpu --extras none > $O/${TAG}_c4_full_solo148.log 2>&1
( time timeout 900 python -m pytest tests/test_gpu_posterior.py -m gpu -q -p no:cacheprovider -k "irt or c3_full or c5_full" ) > $O/${TAG}_pytest_posterior.log 2>&1
echo "pytest rc=$?" >> $O/${TAG}_pytest_posterior.log
ls -la $O | tail -8
