#!/bin/bash
mkdir -p gpurun_out
{
echo "== alt pack x4"; timeout 200 python scripts/dbg_pack.py 4
timeout 600 python -m pytest tests -m gpu -x -q -k "tile_fused or full_size or hodge_blocks" 2>&1 | tail -5
timeout 200 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu 2>/dev/null | cut -c1-330
FQ_TILE_KERNEL=s timeout 200 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu 2>/dev/null | cut -c1-330
} 2>&1 | tee gpurun_out/dbg_pack.log
