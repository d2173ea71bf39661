#!/usr/bin/env bash
# Round 2, fifth hardware session: GLM v5 (two out-of-phase warp groups) vs v2, CTA-per-chain advance step at c4.
set -u
TAG="${1:-r02e}"
O=gpurun_out
mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "glm or irt_latency" ) > $O/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> $O/${TAG}_pytest.log
for V in 1 4; do
  B200NUTS_GLM_VARIANT=$V timeout 240 python bench.py --num-warmup 150 --steps 60 --no-cpu --extras none > $O/${TAG}_c3_glm$V.log 2>&1
  B200NUTS_GLM_VARIANT=$V timeout 240 python bench.py --workload c5 --scale 0.25 --num-warmup 100 --steps 16 --no-cpu --extras none > $O/${TAG}_c5q_glm$V.log 2>&1
done
for A in 0 1; do
  B200NUTS_VERBOSE=1 B200_BENCH_VERBOSE=1 B200NUTS_ADVANCE_CTA=$A timeout 400 python bench.py --workload c4 --no-cpu --extras none > $O/${TAG}_c4_full_cta$A.log 2>&1
done
ls -la $O | tail -8
