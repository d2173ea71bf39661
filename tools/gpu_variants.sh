#!/usr/bin/env bash
# A/B of the GLM kernel variants in one box session: DMMA v2 (B200NUTS_GLM_VARIANT=1), v3 (=2), experimental v4 (=3),
# plus their parity tests and one full ncu capture of the newest kernel.
# Run as: gpurun --timeout 1200 -- bash tools/gpu_variants.sh [tag]
set -u
TAG="${1:-r01d}"
O=gpurun_out
mkdir -p $O
( time B200_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "variants or straggler or bit_identical" ) > $O/${TAG}_pytest_new.log 2>&1
echo "pytest rc=$?" >> $O/${TAG}_pytest_new.log
for V in 1 2 3; do
  B200NUTS_GLM_VARIANT=$V timeout 300 python bench.py --num-warmup 150 --steps 60 --no-cpu > $O/${TAG}_c3_glm$V.log 2>&1
  B200NUTS_GLM_VARIANT=$V timeout 300 python bench.py --workload c5 --scale 0.25 --num-warmup 100 --steps 16 --no-cpu > $O/${TAG}_c5q_glm$V.log 2>&1
done
# c4 straggler path (experimental): host-driven steps vs one CTA per remaining chain running to the end on the device
for T in nf 1; do
  B200_BENCH_VERBOSE=1 B200NUTS_IRT_TAIL=$T timeout 300 python bench.py --workload c4 --chains 1024 --num-warmup 100 --steps 25 --no-cpu > $O/${TAG}_c4_tail_$T.log 2>&1
done
B200NUTS_GLM_VARIANT=3 timeout 400 ncu --set full --clock-control none --import-source on -k regex:glm_dmma4 -s 40 -c 1 -o $O/${TAG}_glm_dmma4 \
   python bench.py --num-warmup 30 --steps 4 --warmup 3 --no-cpu > $O/${TAG}_ncu_full_stdout.log 2>&1
ls -la $O | tail -15
