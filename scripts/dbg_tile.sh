#!/bin/bash
# per-phase warp-cycle statistics of the tile kernel for a list of FQ_TILE_DEBUG values
for dbg in "$@"; do
  export FQ_TILE_DEBUG=$dbg
  echo -n "debug=$dbg  "
  timeout 150 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu 2>/tmp/err.log | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print(round(d['kernels_ms_per_step'].get('k13_tile_fused'),3), end='  ')"
  grep "tile stats" /tmp/err.log | tail -1
done
