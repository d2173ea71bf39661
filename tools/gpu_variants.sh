#!/usr/bin/env bash
# A/B of kernel variants in one box session: GLM DMMA v2 (B200NUTS_GLM_VARIANT=1) vs v3 (=2), IRT one vs two CTAs per SM,
# plus the parity tests of the new code paths and one full ncu capture of the v3 kernel.
# Run as: gpurun --timeout 1200 -- bash tools/gpu_variants.sh [tag]
set -u
TAG="${1:-r01d}"
O=gpurun_out
mkdir -p $O
( time timeout 600 python -m pytest tests/test_gpu_plugin.py tests/test_gpu_parity.py -m gpu -x -q -k "plugin or variants" ) > $O/${TAG}_pytest_new.log 2>&1
echo "pytest rc=$?" >> $O/${TAG}_pytest_new.log
for V in 1 2; do
  B200NUTS_GLM_VARIANT=$V timeout 300 python bench.py --num-warmup 150 --steps 60 --no-cpu > $O/${TAG}_c3_glm$V.log 2>&1
  B200NUTS_GLM_VARIANT=$V timeout 300 python bench.py --workload c5 --scale 0.25 --num-warmup 100 --steps 16 --no-cpu > $O/${TAG}_c5q_glm$V.log 2>&1
  B200NUTS_IRT_VARIANT=$V timeout 300 python bench.py --workload c4 --chains 2048 --num-warmup 60 --steps 10 --no-cpu > $O/${TAG}_c4_irt$V.log 2>&1
done
B200NUTS_GLM_VARIANT=2 timeout 400 ncu --set full --clock-control none --import-source on -k regex:glm_dmma3 -s 40 -c 1 -o $O/${TAG}_glm_dmma3 \
   python bench.py --num-warmup 30 --steps 4 --warmup 3 --no-cpu > $O/${TAG}_ncu_full_stdout.log 2>&1
B200NUTS_IRT_VARIANT=2 timeout 400 ncu --set full --clock-control none --import-source on -k regex:irt_grad -s 40 -c 1 -o $O/${TAG}_irt2 \
   python bench.py --workload c4 --chains 2048 --num-warmup 20 --steps 4 --warmup 3 --no-cpu > $O/${TAG}_ncu_irt_stdout.log 2>&1
ls -la $O | tail -15
