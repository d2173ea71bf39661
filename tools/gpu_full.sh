#!/usr/bin/env bash
# The round's check-point run (what the driver does at round end, on the code as committed): the whole GPU suite, smoke(), the
# default bench line (N = 1: c3 headline + cpu_baseline + the attached c4 / c5 configurations) and the reference arm.
# Run as: gpurun --timeout 3000 -- bash tools/gpu_full.sh <tag>
set -u
TAG="${1:-r02full}"
O=gpurun_out
mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider ) > $O/${TAG}_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/${TAG}_pytest_gpu.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > $O/${TAG}_smoke.log 2>&1
echo "smoke rc=$?" >> $O/${TAG}_smoke.log
( time timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err
( time B200_BENCH_VERBOSE=1 timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 ) > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
echo "bench rc=$?" >> $O/${TAG}_bench.err
ls -la $O | tail -6
