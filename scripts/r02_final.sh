#!/bin/bash
# round-2 final evidence: GPU suite, smoke, the driver's bench line, the reference arm, side workloads, config 3,
# the launch list of the bench command and the ncu full set of the SpMV kernel
mkdir -p gpurun_out
echo "== suite"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r02_final_tests.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "== bench"; timeout 900 python bench.py > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err; tail -c 600 gpurun_out/r02_final_bench.json; tail -3 gpurun_out/r02_final_bench.err
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 0 2>/dev/null | tail -1 > gpurun_out/r02_final_reference.json; cut -c1-300 gpurun_out/r02_final_reference.json
echo "== side workloads"
for w in spmv krylov evp; do timeout 600 python bench.py --workload $w --steps 20 --warmup 3 2>/dev/null | tail -1 | tee -a gpurun_out/r02_final_side.jsonl | cut -c1-400; done
echo "== config 3"; timeout 600 python scripts/config3_4d.py 2>/dev/null | tail -1 | tee gpurun_out/r02_final_config3.json | cut -c1-300
echo "== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_final_launches.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-strong > gpurun_out/r02_final_ncu_launches.log 2>&1; tail -1 gpurun_out/r02_final_ncu_launches.log | cut -c1-200
echo "== ncu spmv"; timeout 600 ncu --set full --import-source on --clock-control none -k regex:stream_reduce_kernel -s 6 -c 1 -o gpurun_out/r02_final_spmv -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-strong --no-kkt > gpurun_out/r02_final_ncu_spmv.log 2>&1; tail -1 gpurun_out/r02_final_ncu_spmv.log | cut -c1-200
