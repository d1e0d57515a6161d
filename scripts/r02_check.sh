#!/bin/bash
# round-2 GPU check: tile tests with the host and the device plan builder, then the whole GPU suite
mkdir -p gpurun_out
echo "== tile tests, host builder"; FQ_TILE_BUILD=host timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tile_fused and not full_size" 2>&1 | tail -5
echo "== whole suite, device builder"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_v5_bench.json 2> gpurun_out/r02_v5_bench.err; tail -c 2500 gpurun_out/r02_v5_bench.json; tail -5 gpurun_out/r02_v5_bench.err
