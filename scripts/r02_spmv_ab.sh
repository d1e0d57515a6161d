#!/bin/bash
# A/B of the SpMV at N=128: each argument is a set of VAR=value assignments
run() { echo "== $*"; env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-kkt 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line); s=d['spmv']; print('fused', round(d['ms_per_step'],3), 'spmv M1 ms', round(s['ms'],4), 'frac', round(s['frac_of_hbm_peak'],3), 'all blocks GB/s', round(s['all_blocks_gbs']))
"; }
for v in "$@"; do run $v; done
