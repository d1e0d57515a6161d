#!/bin/bash
# tuning sweep of the tile-fused kernel on the N=128 workload: "threads brick sig core [debug]"
mkdir -p gpurun_out
run() {
  echo -n "== threads=$1 brick=$2 sig=$3 core=$4 debug=${5:-0}  "
  export FQ_TILE_THREADS=$1 FQ_TILE_SIG=$3 FQ_TILE_CORE=$4 FQ_TILE_DEBUG=${5:-0}
  if [ "$2" != "auto" ]; then export FQ_TILE_BRICK=$2; else unset FQ_TILE_BRICK; fi
  timeout 150 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print(round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['kernels_ms_per_step'].items() if v}, 'frac', round(d['roofline']['frac'],3))
"
}
while read -r line; do [ -n "$line" ] && run $line; done <<LIST
$TUNE_LIST
LIST
