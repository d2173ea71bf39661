#!/usr/bin/env bash
# Round 2, session n: first hardware run of the INT8 digit-plane GLM kernel (variant 7): parity against the oracle, then c3 A/B.
set -u
TAG="${1:-r02n}"
O=gpurun_out
mkdir -p $O
timeout 300 python tools/gpu_i8_check.py > $O/${TAG}_i8_check.txt 2>&1
echo "rc=$?" >> $O/${TAG}_i8_check.txt
tail -20 $O/${TAG}_i8_check.txt
if grep -q "variant 7 N=100000" $O/${TAG}_i8_check.txt; then
  for V in 7 1; do
    B200NUTS_GLM_VARIANT=$V timeout 240 python bench.py --num-warmup 150 --steps 60 --no-cpu --extras none > $O/${TAG}_c3_glm$V.log 2>&1
    tail -c 1500 $O/${TAG}_c3_glm$V.log
  done
fi
