#!/usr/bin/env bash
# Round 2, session u: the INT8 digit-plane kernel as the default for the logistic link -- every GPU test that touches the GLM path
# (gates 1 - 3 incl. the full-size c3 posterior gate), then the c3 bench line with the new roofline fields.
set -u
TAG="${1:-r02u}"
O=gpurun_out
mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_posterior.py tests/test_gpu_cli.py -m gpu -q -p no:cacheprovider -k "glm or log_prob_grad or lockstep or independent or logistic or c3 or cli" ) > $O/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> $O/${TAG}_pytest.log
tail -15 $O/${TAG}_pytest.log
( time timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --extras none ) > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
cut -c 1-400 $O/${TAG}_bench.json
