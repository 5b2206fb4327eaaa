#!/bin/bash
# build SpMV tuning variants: scs_b200/variants/libscs_b200_<tag>.so  (tag:threads:tile:stages:ctas)
cd "$(dirname "$0")/../scs_b200/csrc" || exit 1
mkdir -p ../variants build
for spec in "$@"; do
  IFS=: read tag thr tile st occ <<< "$spec"
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O2 \
       -DSPMV_THREADS=$thr -DSPMV_TILE_NNZ=$tile -DSPMV_STAGES=$st -DSPMV_CTAS_PER_SM=$occ \
       -Xptxas -v -c kernels/spmv.cu -o build/spmv_$tag.o 2> build/spmv_$tag.log || { cat build/spmv_$tag.log; exit 1; }
  grep -h "registers" build/spmv_$tag.log | sort | uniq -c | head -3
  objs=$(ls build/*.cu.o build/*.c.o | grep -v "build/spmv.cu.o")
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/libscs_b200_$tag.so build/spmv_$tag.o $objs \
       -L/usr/local/cuda/lib64 -lcusolver -lcublas -lm
done
ls -la ../variants
