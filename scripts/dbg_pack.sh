#!/bin/bash
mkdir -p gpurun_out
{
echo "== alt nopack x6"; FQ_TILE_PACK=0 timeout 200 python scripts/dbg_pack.py 6
echo "== alt pack x6"; timeout 200 python scripts/dbg_pack.py 6
echo "== alt pack regs0 x6"; FQ_ALT_REGS=0 timeout 200 python scripts/dbg_pack.py 6
timeout 600 python -m pytest tests -m gpu -x -q -k "tile_fused or full_size" 2>&1 | tail -5
timeout 200 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu 2>/dev/null | cut -c1-400
} 2>&1 | tee gpurun_out/dbg_pack.log
