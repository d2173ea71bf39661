#!/usr/bin/env bash
# quick GPU check: parity tests + c3 bench without the CPU leg (+ optional ncu full capture of the GLM kernel)
set -u
TAG="${1:-q}"
O=gpurun_out
mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/${TAG}_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/${TAG}_pytest_gpu.log
tail -5 $O/${TAG}_pytest_gpu.log
( time timeout 900 python bench.py --no-cpu ) > $O/${TAG}_bench_c3.log 2>&1
tail -4 $O/${TAG}_bench_c3.log
( time timeout 600 python bench.py --workload c2 --no-cpu ) > $O/${TAG}_bench_c2.log 2>&1
tail -4 $O/${TAG}_bench_c2.log
if [ "${2:-}" = "ncu" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:glm_dmma -s 40 -c 1 -o $O/${TAG}_glm_dmma \
   python bench.py --num-warmup 30 --steps 4 --warmup 3 --no-cpu > $O/${TAG}_ncu_full_stdout.log 2>&1
fi
