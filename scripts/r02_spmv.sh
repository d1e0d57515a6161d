#!/bin/bash
# SpMV: parity tests, then the bench's SpMV numbers (M1 and the KKT operator) at N=128
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "spmv or cg or minres or lanczos or window or jacobi or kkt" 2>&1 | tail -4
timeout 400 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line); print(d['ms_per_step']); print(json.dumps(d['spmv']))
"
