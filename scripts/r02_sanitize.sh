#!/bin/bash
# compute-sanitizer memcheck over the small tile / SpMV / Krylov tests (no full-size cases)
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/r02_memcheck.log \
  python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "(tile_fused and not full_size and not many_tiles) or spmv_bitwise or device_resident or partitioned_uploaded or async_download" 2>&1 | tail -4
echo "memcheck rc=$?"; grep -c 'ERROR SUMMARY' gpurun_out/r02_memcheck.log; tail -3 gpurun_out/r02_memcheck.log
