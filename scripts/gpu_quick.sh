#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_linsys_gpu.py tests/test_solver_gpu.py -x -q -m gpu 2>&1 | tail -3
REPS=20 python scripts/prof_spmv.py 2>&1 | tail -3
TAG=${TAG:-r01c}
ncu --set full --clock-control none --import-source on -k regex:spmv_ -s 8 -c 4 \
    -o gpurun_out/spmv_${TAG} -f python scripts/prof_spmv.py > gpurun_out/ncu_spmv_${TAG}.log 2>&1
ls -la gpurun_out | tail -4
