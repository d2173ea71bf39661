#!/usr/bin/env bash
# A/B of the GLM kernel variants in one box session: DMMA v2 (B200NUTS_GLM_VARIANT=1), v3 (=2), experimental v4 (=3),
# plus their parity tests and one full ncu capture of the newest kernel.
# Run as: gpurun --timeout 1200 -- bash tools/gpu_variants.sh [tag]
set -u
TAG="${1:-r02a}"
O=gpurun_out
mkdir -p $O
# the whole GPU suite, experimental kernels included, no -x: the tests written after the last hardware session of round 1
# (tests/conftest.py NOT_YET_RUN_ON_HARDWARE) are collected last
( time B200_TEST_EXPERIMENTAL=1 timeout 900 python -m pytest tests -m gpu -q ) > $O/${TAG}_pytest_gpu_all.log 2>&1
echo "pytest rc=$?" >> $O/${TAG}_pytest_gpu_all.log
for V in 1 2 3; do
  B200NUTS_GLM_VARIANT=$V timeout 300 python bench.py --num-warmup 150 --steps 60 --no-cpu > $O/${TAG}_c3_glm$V.log 2>&1
  B200NUTS_GLM_VARIANT=$V timeout 300 python bench.py --workload c5 --scale 0.25 --num-warmup 100 --steps 16 --no-cpu > $O/${TAG}_c5q_glm$V.log 2>&1
done
# c4 straggler path (experimental): host-driven steps vs one CTA per remaining chain running to the end on the device
for T in nf 1; do
  B200_BENCH_VERBOSE=1 B200NUTS_IRT_TAIL=$T timeout 300 python bench.py --workload c4 --chains 1024 --num-warmup 100 --steps 25 --no-cpu > $O/${TAG}_c4_tail_$T.log 2>&1
done
# the full c4 configuration with the straggler kernel (r01h without it: 11.4 s for 30 transitions, 1 598 ESS/s)
B200_BENCH_VERBOSE=1 B200NUTS_IRT_TAIL=1 timeout 420 python bench.py --workload c4 --no-cpu > $O/${TAG}_c4_full_tail.log 2>&1
# c2 on the final arrangement of the chain kernels (r01j: 1.04e7 ESS/s)
timeout 300 python bench.py --workload c2 --steps 200 > $O/${TAG}_bench_c2.log 2>&1
B200NUTS_GLM_VARIANT=3 timeout 400 ncu --set full --clock-control none --import-source on -k regex:glm_dmma4 -s 40 -c 1 -o $O/${TAG}_glm_dmma4 \
   python bench.py --num-warmup 30 --steps 4 --warmup 3 --no-cpu > $O/${TAG}_ncu_full_stdout.log 2>&1
ls -la $O | tail -15
