#!/bin/bash
# A/B of the tile kernels on the N=128 workload (gpurun): parity tests of the warp-specialised variants first,
# then ms/step of default / pipelined (p) / old warp-specialised (w), and p with one side disabled.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "tile_fused" > gpurun_out/exp_ws2_tests.log 2>&1
echo "tests exit $?" | tee -a gpurun_out/exp_ws2_tests.log
tail -3 gpurun_out/exp_ws2_tests.log
run() {
  echo -n "== kernel='$1' debug=$2 brick=${3:-auto}  "
  export FQ_TILE_KERNEL=$1 FQ_TILE_DEBUG=$2
  if [ -n "$3" ]; then export FQ_TILE_BRICK=$3; else unset FQ_TILE_BRICK; fi
  timeout 200 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu 2>gpurun_out/exp_ws2_err.log | python -c "
import json,sys
l=sys.stdin.readline()
try:
    d=json.loads(l)
    print(round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['kernels_ms_per_step'].items() if v}, 'frac', round(d['roofline']['frac'],3))
except Exception as e:
    print('FAILED', l[:200])
"
}
{
FQ_TILE_PACK_STATS=1 run "" 0
grep "tile pack" gpurun_out/exp_ws2_err.log
}
export FQ_TILE_KERNEL= FQ_TILE_DEBUG=0
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tile_assemble -s 2 -c 1 -f -o gpurun_out/r01_tile_pack_n128 \
    python scripts/microbench.py --n 128 --reps 1 --fused > gpurun_out/r01_tile_pack_n128.log 2>&1
{
tail -2 gpurun_out/r01_tile_pack_n128.log
} 2>&1 | tee gpurun_out/exp_ws2.log
