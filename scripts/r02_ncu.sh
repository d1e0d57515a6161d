#!/bin/bash
# ncu full-set capture of the fused kernel at N=128 (one launch), plus the launch list of the bench command
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:tile_fused_kernel -s 3 -c 1 -o gpurun_out/r02_fused_full -f env FQ_TILE_WARPS=${FQ_TILE_WARPS:-0} python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/r02_ncu_full.log 2>&1
tail -3 gpurun_out/r02_ncu_full.log
