#!/usr/bin/env bash
# Round 2, session k: GLM v7 (one warp per sub-partition, link woven into phase B) vs v2.
set -u
TAG="${1:-r02k}"
O=gpurun_out
mkdir -p $O
( time timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "glm" ) > $O/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> $O/${TAG}_pytest.log
for V in 1 6; do
  B200NUTS_GLM_VARIANT=$V timeout 240 python bench.py --num-warmup 150 --steps 60 --no-cpu --extras none > $O/${TAG}_c3_glm$V.log 2>&1
  B200NUTS_GLM_VARIANT=$V timeout 240 python bench.py --workload c5 --scale 0.25 --num-warmup 100 --steps 16 --no-cpu --extras none > $O/${TAG}_c5q_glm$V.log 2>&1
done
B200NUTS_GLM_VARIANT=6 timeout 400 ncu --set full --clock-control none --import-source on -k regex:glm_dmma7 -s 40 -c 1 -o $O/${TAG}_glm7 \
   python bench.py --num-warmup 30 --steps 4 --warmup 3 --no-cpu --extras none > $O/${TAG}_ncu_glm7_stdout.log 2>&1
ls -la $O | tail -6
