#!/usr/bin/env bash
# One GPU-box session, most important evidence first: parity tests, headline bench (+ warm-start fixture for the
# reference arm), ncu launch list and full capture of the dominant kernel, then the other single-GPU workloads.
# Run as: gpurun --timeout 2400 -- bash tools/gpu_round.sh [tag]
set -u
TAG="${1:-r01}"
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.txt 2>&1
nproc > $O/${TAG}_nproc.txt
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/${TAG}_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/${TAG}_pytest_gpu.log
( time timeout 600 python bench.py --dump-warm $O/warm_c3.npz ) > $O/${TAG}_bench_c3.log 2>&1
B200_WARM_FIXTURE_DIR=$O timeout 400 python bench.py --impl reference --steps 20 --warmup 3 > $O/${TAG}_bench_c3_reference.log 2>&1
# launch list of a short run of the same command (shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file $O/${TAG}_launches_c3.csv \
   python bench.py --num-warmup 30 --steps 4 --warmup 3 --no-cpu > $O/${TAG}_ncu_launches_stdout.log 2>&1
# full capture of the dominant kernel
timeout 600 ncu --set full --clock-control none --import-source on -k regex:glm_dmma -s 40 -c 1 -o $O/${TAG}_glm_dmma \
   python bench.py --num-warmup 30 --steps 4 --warmup 3 --no-cpu > $O/${TAG}_ncu_full_stdout.log 2>&1
( time timeout 300 python bench.py --workload c2 --steps 200 ) > $O/${TAG}_bench_c2.log 2>&1
( time timeout 300 python bench.py --workload c1 --steps 1000 ) > $O/${TAG}_bench_c1.log 2>&1
( time B200_BENCH_VERBOSE=1 timeout 420 python bench.py --workload c4 --no-cpu ) > $O/${TAG}_bench_c4.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:irt_grad -s 40 -c 1 -o $O/${TAG}_irt \
   python bench.py --workload c4 --chains 2048 --num-warmup 20 --steps 4 --warmup 3 --no-cpu > $O/${TAG}_ncu_irt_stdout.log 2>&1
( time timeout 420 python bench.py --workload c5 --no-cpu ) > $O/${TAG}_bench_c5.log 2>&1
ls -la $O
