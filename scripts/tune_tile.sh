#!/bin/bash
# tuning sweep of the tile-fused kernel on the N=128 workload: "threads brick sig stages"
mkdir -p gpurun_out
run() {
  echo "== threads=$1 brick=$2 sig=$3 stages=${4:-default}"
  export FQ_TILE_THREADS=$1 FQ_TILE_SIG=$3
  if [ "$2" != "auto" ]; then export FQ_TILE_BRICK=$2; else unset FQ_TILE_BRICK; fi
  if [ -n "$4" ]; then export FQ_TILE_STAGES=$4; else unset FQ_TILE_STAGES; fi
  python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print(json.dumps({k:d[k] for k in ('value','ms_per_step','kernels_ms_per_step','symbolic_ms')}), 'frac', d['roofline']['frac'])
"
}
while read -r line; do [ -n "$line" ] && run $line; done <<LIST
$TUNE_LIST
LIST
