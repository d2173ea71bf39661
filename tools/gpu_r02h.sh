#!/usr/bin/env bash
# Round 2, session h: DMMA issue-rate microbenchmark, GLM v6 (16 warps in two out-of-phase groups) vs v2, ncu of the CTA advance step.
set -u
TAG="${1:-r02h}"
O=gpurun_out
mkdir -p $O
timeout 120 tools/microbench/dmma_shapes > $O/${TAG}_dmma_shapes.txt 2>&1
( time timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "glm" ) > $O/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> $O/${TAG}_pytest.log
for V in 1 5; do
  B200NUTS_GLM_VARIANT=$V timeout 240 python bench.py --num-warmup 150 --steps 60 --no-cpu --extras none > $O/${TAG}_c3_glm$V.log 2>&1
  B200NUTS_GLM_VARIANT=$V timeout 240 python bench.py --workload c5 --scale 0.25 --num-warmup 100 --steps 16 --no-cpu --extras none > $O/${TAG}_c5q_glm$V.log 2>&1
done
B200NUTS_GLM_VARIANT=5 timeout 400 ncu --set full --clock-control none --import-source on -k regex:glm_dmma6 -s 40 -c 1 -o $O/${TAG}_glm6 \
   python bench.py --num-warmup 30 --steps 4 --warmup 3 --no-cpu --extras none > $O/${TAG}_ncu_glm6_stdout.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:advance_cta -s 300 -c 1 -o $O/${TAG}_advcta \
   python bench.py --workload c4 --num-warmup 10 --steps 4 --warmup 3 --no-cpu --extras none > $O/${TAG}_ncu_advcta_stdout.log 2>&1
ls -la $O | tail -8
