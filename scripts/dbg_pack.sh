#!/bin/bash
mkdir -p gpurun_out
{
for nc in 20 24; do
echo "== consumers $nc"; FQ_ALT_CONSUMERS=$nc timeout 200 python scripts/dbg_pack.py 2
FQ_ALT_CONSUMERS=$nc timeout 200 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu 2>/dev/null | cut -c1-260
done
timeout 200 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu 2>/dev/null | cut -c1-260
} 2>&1 | tee gpurun_out/dbg_pack.log
