#!/usr/bin/env bash
# Round 2, sixth hardware session: where a c4 lock-step goes (launch list), IRT throughput form with packed responses (1 vs 2 CTAs per SM).
set -u
TAG="${1:-r02f}"
O=gpurun_out
mkdir -p $O
for A in 1 0; do
  B200NUTS_ADVANCE_CTA=$A timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 400 --csv --log-file $O/${TAG}_launches_c4_cta$A.csv \
     python bench.py --workload c4 --num-warmup 20 --steps 4 --warmup 3 --no-cpu --extras none > $O/${TAG}_ncu_c4_cta$A.log 2>&1
done
for V in 1 2; do
  B200NUTS_VERBOSE=1 B200_BENCH_VERBOSE=1 B200NUTS_IRT_VARIANT=$V timeout 400 python bench.py --workload c4 --no-cpu --extras none > $O/${TAG}_c4_full_irt$V.log 2>&1
done
( time timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_posterior.py -m gpu -q -p no:cacheprovider -k "irt" ) > $O/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> $O/${TAG}_pytest.log
ls -la $O | tail -8
