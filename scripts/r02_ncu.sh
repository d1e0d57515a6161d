#!/bin/bash
# round-2 evidence: ncu full-set capture of the fused kernel at N=128 (one launch) + the launch list of the bench command
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:tile_fused_kernel -s 4 -c 1 -o gpurun_out/r02_fused_full -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-strong --no-kkt > gpurun_out/r02_ncu_full.log 2>&1
tail -2 gpurun_out/r02_ncu_full.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_n128.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-strong > gpurun_out/r02_ncu_launches.log 2>&1
tail -2 gpurun_out/r02_ncu_launches.log
timeout 600 ncu --set full --clock-control none -k regex:stream_reduce_kernel -s 6 -c 1 -o gpurun_out/r02_spmv_full -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-strong --no-kkt > gpurun_out/r02_ncu_spmv.log 2>&1
tail -2 gpurun_out/r02_ncu_spmv.log
