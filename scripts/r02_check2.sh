#!/bin/bash
mkdir -p gpurun_out
echo "== new tests"; timeout 1500 python -m pytest tests -m gpu -x -q -k "hdif_gram or afw or lanczos or degenerate or c_program or eigen" 2>&1 | tail -15
echo "== evp 1 rank"; timeout 600 python scripts/evp_multi_gpu.py --grid 6 2>&1 | tail -3
