#!/bin/bash
# compute-sanitizer memcheck over the GPU tests (without the full-size / long-running cases); $1 = pytest -k expression
mkdir -p gpurun_out
K=${1:-"(tile_fused and not full_size and not many_tiles) or spmv_bitwise or device_resident or partitioned_uploaded or async_download"}
timeout 2400 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/r02_memcheck.log \
  python -m pytest tests/test_gpu_parity.py -m gpu -q -k "$K" 2>&1 | tail -4
tail -2 gpurun_out/r02_memcheck.log
