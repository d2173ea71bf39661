#!/usr/bin/env bash
set -u
TAG="${1:-r01j}"
O=gpurun_out
mkdir -p $O
( time timeout 100 python bench.py --workload c2 --steps 200 --no-cpu ) > $O/${TAG}_bench_c2.log 2>&1
( time timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cli.py -m gpu -q -k "dense or advance_kernel or engines_identical or matched_stream or independent_of_batch or sibling" ) > $O/${TAG}_pytest_subset.log 2>&1
echo "pytest rc=$?" >> $O/${TAG}_pytest_subset.log
