#!/usr/bin/env bash
# Round 2, session i: two-kernel advance step for long parameter vectors, IRT throughput form with three CTAs per SM.
set -u
TAG="${1:-r02i}"
O=gpurun_out
mkdir -p $O
( time timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "irt" ) > $O/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> $O/${TAG}_pytest.log
for cfg in "1 2" "1 3" "0 2"; do
  set -- $cfg
  B200NUTS_VERBOSE=1 B200_BENCH_VERBOSE=1 B200NUTS_ADVANCE_CTA=$1 B200NUTS_IRT_VARIANT=$2 timeout 400 python bench.py --workload c4 --no-cpu --extras none > $O/${TAG}_c4_cta$1_irt$2.log 2>&1
done
B200NUTS_IRT_VARIANT=3 timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 300 --csv --log-file $O/${TAG}_launches_c4.csv \
   python bench.py --workload c4 --num-warmup 20 --steps 4 --warmup 3 --no-cpu --extras none > $O/${TAG}_ncu_c4.log 2>&1
ls -la $O | tail -6
