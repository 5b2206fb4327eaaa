#!/bin/bash
# v3 (flagged stream) SpMV tuning variants: tag:threads:stages:ctas[:extra -D flags]
cd "$(dirname "$0")/../scs_b200/csrc" || exit 1
mkdir -p ../variants build
for spec in "$@"; do
  IFS=: read tag thr st occ extra <<< "$spec"
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O2 \
       -DSPMV3_THREADS=$thr -DSPMV3_STAGES=$st -DSPMV3_CTAS_PER_SM=$occ $extra \
       -Xptxas -v -c kernels/spmv.cu -o build/spmv_$tag.o 2> build/spmv_$tag.log || { cat build/spmv_$tag.log; exit 1; }
  echo "$tag: $(grep -A1 'spmv_flag_kernelILi0' build/spmv_$tag.log | grep -o 'Used [0-9]* registers' | head -1) $(grep -A1 'spmv_flag_kernelILi0' build/spmv_$tag.log | grep -o '[0-9]* bytes spill stores' | head -1)"
  objs=$(ls build/*.cu.o build/*.c.o | grep -v "build/spmv.cu.o")
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/libscs_b200_$tag.so build/spmv_$tag.o $objs \
       -L/usr/local/cuda/lib64 -lcusolver -lcublas -lm
done
ls ../variants
