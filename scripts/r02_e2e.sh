#!/bin/bash
# e2e leg of the bench; each argument is a set of VAR=value assignments
for v in "$@"; do
  echo "== $v"
  env $v FQ_BENCH_VERBOSE=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-kkt 2> gpurun_out/e2e.err | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line); e=d['e2e']; print(round(e['ms_per_step'],1), 'ms', e['h2d_bytes_per_step'], e['d2h_pcie_bytes_per_step'])
"
  grep 'e2e step' gpurun_out/e2e.err
done
