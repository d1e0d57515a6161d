#!/bin/bash
# round-2 GPU check: the whole GPU suite, then the default bench line
mkdir -p gpurun_out
echo "== whole suite"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_v5_bench.json 2> gpurun_out/r02_v5_bench.err; tail -c 1800 gpurun_out/r02_v5_bench.json; tail -5 gpurun_out/r02_v5_bench.err
