#!/bin/bash
# round-1 evidence for the default tile kernel (alternating producer/consumer kernel, bank-aware packed streams):
# launch list of the default bench command, full ncu set of the tile kernel at N=128, traffic json, bench lines.
mkdir -p gpurun_out
# 1. every launch of the default bench command (symbolic phase, first numeric pass, plan build, then the timed steps =
#    one tile kernel launch each, then the SpMV section)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_launches_bench_n128.csv \
    python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/r01_launches_bench.log 2>&1
# 2. full set + source counters of one steady-state launch
ncu --set full --clock-control none --import-source on -k regex:tile_assemble -s 2 -c 1 -f -o gpurun_out/r01_tile_alt_n128 \
    python scripts/microbench.py --n 128 --reps 1 --fused > gpurun_out/r01_tile_alt_n128.log 2>&1
ncu -i gpurun_out/r01_tile_alt_n128.ncu-rep --page raw --csv > gpurun_out/r01_tile_alt_ncu_raw_n128.csv 2>/dev/null
ncu -i gpurun_out/r01_tile_alt_n128.ncu-rep --page details > gpurun_out/r01_tile_alt_ncu_details_n128.txt 2>/dev/null
python scripts/ncu_summary.py gpurun_out/r01_tile_alt_n128.ncu-rep 12582912 > gpurun_out/r01_tile_alt_ncu_summary_n128.txt 2>&1
python - <<'PY'
import csv, json
rows = list(csv.reader(open("gpurun_out/r01_tile_alt_ncu_raw_n128.csv")))
hdr, units, vals = rows[0], rows[1], rows[2]
d, u = dict(zip(hdr, vals)), dict(zip(hdr, units))
def b(k):
    return float(d[k].replace(",", "")) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[u[k]]
r, w = b("dram__bytes_read.sum"), b("dram__bytes_write.sum")
out = {"n": 128, "cells": 12582912, "kernel": d["Kernel Name"], "dram_bytes_read": r, "dram_bytes_write": w,
       "dram_bytes_per_launch": r + w, "gpu_time_ms_under_ncu": float(d["gpu__time_duration.sum"]),
       "source": "profiles/r01_tile_alt_ncu_raw_n128.csv (ncu --set full, one launch, N=128)"}
json.dump(out, open("gpurun_out/r01_tile_traffic.json", "w"), indent=1)
json.dump(out, open("profiles/r01_tile_traffic.json", "w"), indent=1)
print(out)
PY
# 3. the bench lines (the second one is the default command)
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01_bench_reference.json 2> gpurun_out/r01_bench_reference.err
python bench.py > gpurun_out/r01_final_bench_n128.json 2> gpurun_out/r01_final_bench_n128.err
tail -c 1800 gpurun_out/r01_final_bench_n128.json
