#!/bin/bash
mkdir -p gpurun_out
{
for v in 3 4 5; do
echo "== FQ_ALT_REGS=$v"; FQ_ALT_REGS=$v timeout 200 python scripts/dbg_pack.py 2
FQ_ALT_REGS=$v timeout 200 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu 2>/dev/null | cut -c1-260
done
timeout 200 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu 2>/dev/null | cut -c1-260
} 2>&1 | tee gpurun_out/dbg_pack.log
