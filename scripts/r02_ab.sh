#!/bin/bash
# A/B timings of the fused kernel at N=128
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line); print(d['ms_per_step'], d['roofline']['kernel_ms_per_launch'])
"; }
for v in "$@"; do run $v; done
