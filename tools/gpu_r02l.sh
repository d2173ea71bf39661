#!/usr/bin/env bash
# Round 2, session l: design matrix uploaded straight into the kernels' layout (e2e host->device leg), the tests changed since the
# check-point run r02j.
set -u
TAG="${1:-r02l}"
O=gpurun_out
mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_cli.py tests/test_gpu_plugin.py tests/test_gpu_posterior.py tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "cli or plugin or irt_vs_reference or lp_grad or glm or full_size" ) > $O/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> $O/${TAG}_pytest.log
( time timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --extras none ) > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
ls -la $O | tail -4
