#!/usr/bin/env bash
# Short box session: parity tests of the new code paths (plug-in models, kernel variants, dense_e) and a c4 diagnosis
# (where does an IRT lock-step go: gradient kernel vs advance kernel), both IRT kernel forms.
set -u
TAG="${1:-r01e}"
O=gpurun_out
mkdir -p $O
( time timeout 700 python -m pytest tests/test_gpu_plugin.py tests/test_gpu_parity.py tests/test_gpu_cli.py -m gpu -q -k "plugin or variants or dense" ) > $O/${TAG}_pytest_new.log 2>&1
echo "pytest rc=$?" >> $O/${TAG}_pytest_new.log
for V in 1 2; do
  B200_BENCH_VERBOSE=1 B200NUTS_IRT_VARIANT=$V timeout 240 python bench.py --workload c4 --chains 1024 --num-warmup 100 --steps 20 --no-cpu > $O/${TAG}_c4_irt$V.log 2>&1
done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 200 --csv --log-file $O/${TAG}_launches_c4.csv \
   python bench.py --workload c4 --chains 1024 --num-warmup 20 --steps 4 --warmup 3 --no-cpu > $O/${TAG}_ncu_launches_c4_stdout.log 2>&1
ls -la $O | tail -8
