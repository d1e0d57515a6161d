#!/bin/bash
export FQ_TILE_THREADS=512 FQ_TILE_BRICK=5,3,2 FQ_TILE_SIG=1 FQ_TILE_STAGES=3
for dbg in 0 1 2 3 4 6 7; do
  export FQ_TILE_DEBUG=$dbg
  echo -n "debug=$dbg  "
  python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print(d['kernels_ms_per_step'].get('k13_tile_fused'))"
done
