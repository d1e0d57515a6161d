#!/bin/bash
# e2e leg of the bench (host-side widening of the index arrays on by default; FQ_HOST_WIDEN_THREADS=0 = device widening)
for t in "$@"; do
  echo "== FQ_HOST_WIDEN_THREADS=$t"
  if [ "$t" != "default" ]; then export FQ_HOST_WIDEN_THREADS=$t; fi
  FQ_BENCH_VERBOSE=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-kkt 2> gpurun_out/e2e_$t.err | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line); print(json.dumps(d['e2e']))
"
  grep 'e2e step' gpurun_out/e2e_$t.err
done
