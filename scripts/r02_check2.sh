#!/bin/bash
# whole GPU suite, then the SpMV / fused numbers of the bench
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
bash scripts/r02_spmv_ab.sh FQ_X=0
