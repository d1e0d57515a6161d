#!/bin/bash
# usage: scripts/build_variant.sh <name> "<extra nvcc flags>" file1 [file2 ...]   (files without .cu, recompiled with the flags)
# links formoniq_b200/lib/libfq_<name>.so from the recompiled objects and the regular ones (select it with FQ_LIB_PATH)
name=$1; flags=$2; shift 2
OBJ=formoniq_b200/build; T=/tmp/fqv_$name; mkdir -p $T
for f in "$@"; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC -ccbin /usr/bin/g++ $flags -c formoniq_b200/csrc/$f.cu -o $T/$f.o >/dev/null 2>&1 &
done
wait
objs=""
for f in elmat kuhn assemble tile blockop matfree quadform spmv blas1 krylov capi; do
  if [ -f $T/$f.o ]; then objs="$objs $T/$f.o"; else objs="$objs $OBJ/$f.o"; fi
done
/usr/local/cuda/bin/nvcc -shared -o formoniq_b200/lib/libfq_$name.so $objs -cudart static -ccbin /usr/bin/g++ -gencode arch=compute_100a,code=sm_100a 2>/dev/null && echo built libfq_$name.so
