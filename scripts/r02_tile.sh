#!/bin/bash
# tile-fused kernel: parity tests, then timings
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tile_fused or hodge or oracle_parity or many_tiles" 2>&1 | tail -3
bash scripts/r02_ab.sh FQ_X=0 FQ_TILE_WARPS=1 FQ_TILE_WARPS=2 FQ_TILE_DEBUG=2
