#!/usr/bin/env bash
# quick iteration on the INT8 GLM kernel: parity against the oracle, role trace of one launch, short c3 bench
set -u
TAG="${1:-i8}"
O=gpurun_out
mkdir -p $O
timeout 300 python tools/gpu_i8_check.py > $O/${TAG}_i8_check.txt 2>&1
grep "variant 7" $O/${TAG}_i8_check.txt | cut -c 1-150
B200NUTS_I8_TRACE=$O/${TAG}_trace.bin timeout 240 python bench.py --num-warmup 150 --steps 60 --no-cpu --extras none > $O/${TAG}_c3.log 2>&1
python tools/i8_trace.py $O/${TAG}_trace.bin 3 | tail -4
python - <<PY
import json
t=open("$O/${TAG}_c3.log").read()
d=json.loads(t[t.index('{"metric"'):].splitlines()[0]); r=d['roofline']
print("value", round(d['value']), "ms/step", round(d['ms_per_step'],3), "avg_launch_ms", round(r['avg_launch_ms'],4), "chains/launch", round(r['avg_chains_per_launch']), "TF", round(r['achieved'],1), "e2e", round(d['e2e']['value']))
PY
