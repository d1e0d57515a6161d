#!/bin/bash
# round-1 evidence: launch list of the default bench command + full ncu set of the tile-fused kernel at N=128
mkdir -p gpurun_out
# every launch of the default bench command (symbolic phase, first numeric pass, plan build, then the timed steps =
# one tile_assemble_kernel launch each, then the SpMV section)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_launches_n128.csv \
    python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/r01_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tile_assemble -s 2 -c 1 -o gpurun_out/r01_tile_n128 \
    python scripts/microbench.py --n 128 --reps 1 --fused > gpurun_out/r01_tile_n128.log 2>&1
tail -3 gpurun_out/r01_tile_n128.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r01_bench_n128.json 2> gpurun_out/r01_bench_n128.err
tail -c 1500 gpurun_out/r01_bench_n128.json
