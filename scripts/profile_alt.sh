#!/bin/bash
# ncu --set full (with source counters) of the alternating tile kernel at N=128
mkdir -p gpurun_out
export FQ_TILE_KERNEL=a FQ_ALT_REGS=${FQ_ALT_REGS:-2}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tile_assemble -s 2 -c 1 -f -o gpurun_out/r01_tile_alt_n128 \
    python scripts/microbench.py --n 128 --reps 1 --fused > gpurun_out/r01_tile_alt_n128.log 2>&1
tail -3 gpurun_out/r01_tile_alt_n128.log
ls -la gpurun_out/
