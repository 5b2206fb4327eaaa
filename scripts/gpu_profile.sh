#!/bin/bash
# ncu evidence for the hot kernels (run under gpurun, ONE GPU). Outputs -> gpurun_out/
mkdir -p gpurun_out
TAG=${TAG:-r01}
python scripts/prof_spmv.py 2>&1 | tail -4
# full section set for the SpMV kernels (skip warm-up launches)
ncu --set full --clock-control none --import-source on -k regex:spmv_ -s 8 -c 4 \
    -o gpurun_out/spmv_${TAG} -f python scripts/prof_spmv.py > gpurun_out/ncu_spmv_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_cg_ -s 4 -c 4 \
    -o gpurun_out/cgvec_${TAG} -f python scripts/prof_spmv.py > gpurun_out/ncu_cgvec_${TAG}.log 2>&1
# launch list of a short solve (shares of the step)
ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1200 --csv \
    --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_${TAG}.log 2>&1
ls -la gpurun_out | tail -12
