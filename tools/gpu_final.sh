#!/usr/bin/env bash
# Final box session of a round on the final code: the whole GPU suite, the headline bench with its CPU leg, the reference
# arm from the refreshed warm-start fixture, and c4 at full size.
set -u
TAG="${1:-r01g}"
O=gpurun_out
mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -q ) > $O/${TAG}_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/${TAG}_pytest_gpu.log
( time timeout 400 python bench.py --dump-warm $O/warm_c3.npz ) > $O/${TAG}_bench_c3.log 2>&1
B200_WARM_FIXTURE_DIR=$O timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > $O/${TAG}_bench_c3_reference.log 2>&1
( time B200_BENCH_VERBOSE=1 timeout 420 python bench.py --workload c4 --num-warmup 150 --steps 30 --no-cpu ) > $O/${TAG}_bench_c4.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1
ls -la $O | tail -8
