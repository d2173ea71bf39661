#!/usr/bin/env bash
# Round 2, final check-point on the code as shipped: the whole GPU suite, smoke(), the bench line exactly as the driver runs it (N = 1) and
# its reference arm, and two full ncu captures of the INT8 kernel: caches flushed between replays (ncu's default: cold-cache traffic) and
# caches left alone (the steady state of a run, where the digit planes of X stay in L2 between launches).
set -u
TAG="${1:-r03j}"
O=gpurun_out
mkdir -p $O
( time timeout 1100 python -m pytest tests -m gpu -q -p no:cacheprovider ) > $O/${TAG}_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/${TAG}_pytest_gpu.log
tail -6 $O/${TAG}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" >> $O/${TAG}_smoke.log
( time timeout 800 python bench.py --gpus 1 --steps 20 --warmup 5 ) > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
( time timeout 400 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err
timeout 300 ncu --set full --clock-control none --cache-control none --import-source on -k regex:glm_i8_kernel -s 40 -c 1 -o $O/${TAG}_glm_i8_warm \
   python bench.py --num-warmup 30 --steps 4 --warmup 3 --no-cpu --extras none > $O/${TAG}_ncu_warm_stdout.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:glm_i8_kernel -s 40 -c 1 -o $O/${TAG}_glm_i8 \
   python bench.py --num-warmup 30 --steps 4 --warmup 3 --no-cpu --extras none > $O/${TAG}_ncu_stdout.log 2>&1
ls -la $O | tail -8
