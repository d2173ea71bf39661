#!/usr/bin/env bash
# Round 2, session o: ncu --set full of the INT8 digit-plane GLM kernel inside the c3 bench.
set -u
TAG="${1:-r02o}"
O=gpurun_out
mkdir -p $O
B200NUTS_GLM_VARIANT=7 timeout 500 ncu --set full --clock-control none --import-source on -k regex:glm_i8_kernel -s 40 -c 1 -o $O/${TAG}_glm_i8 \
   python bench.py --num-warmup 30 --steps 4 --warmup 3 --no-cpu --extras none > $O/${TAG}_ncu_glm_i8_stdout.log 2>&1
tail -3 $O/${TAG}_ncu_glm_i8_stdout.log | cut -c 1-600
ls -la $O | tail -3
