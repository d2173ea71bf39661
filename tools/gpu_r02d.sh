#!/usr/bin/env bash
# Round 2, fourth hardware session: ncu captures of the two dominant kernels as shipped (IRT throughput form, GLM DMMA v2 with the
# product-form link), the c4 phase split, the plug-in / multi-rank-free tests added since r02c.
set -u
TAG="${1:-r02d}"
O=gpurun_out
mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:irt_grad_kernel -s 60 -c 1 -o $O/${TAG}_irt \
   python bench.py --workload c4 --chains 2368 --num-warmup 10 --steps 4 --warmup 3 --no-cpu --extras none > $O/${TAG}_ncu_irt_stdout.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:glm_dmma_kernel -s 40 -c 1 -o $O/${TAG}_glm \
   python bench.py --num-warmup 30 --steps 4 --warmup 3 --no-cpu --extras none > $O/${TAG}_ncu_glm_stdout.log 2>&1
B200NUTS_VERBOSE=1 B200_BENCH_VERBOSE=1 timeout 400 python bench.py --workload c4 --no-cpu --extras none > $O/${TAG}_c4_full.log 2>&1
( time timeout 600 python -m pytest tests/test_gpu_plugin.py tests/test_gpu_posterior.py -m gpu -q -p no:cacheprovider -k "plugin or c4_full or irt" ) > $O/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> $O/${TAG}_pytest.log
ls -la $O | tail -8
