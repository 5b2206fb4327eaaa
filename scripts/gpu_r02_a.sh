#!/bin/bash
# round 2, call A: scatter/one-pass ceiling, staged tests, fused K3+K4, full suite
mkdir -p gpurun_out
L=gpurun_out/r02a.log
: > $L
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv >> $L
echo "=== scatter bench" >> $L
timeout 300 scripts/bin/scatter_bench >> $L 2>&1
echo "=== staged tests (-m gpu_unverified)" >> $L
timeout 900 python -m pytest tests -q -m gpu_unverified -s 2>&1 | tail -60 >> $L
echo "=== staged: K3+K4 fused (SCS_B200_FUSE_K34=1): linsys + solver tests, CG iteration timing" >> $L
SCS_B200_FUSE_K34=1 timeout 600 python -m pytest tests/test_linsys_gpu.py tests/test_solver_gpu.py -x -q -m gpu 2>&1 | tail -5 >> $L
REPS=20 timeout 300 python scripts/prof_spmv.py 2>&1 | tail -3 >> $L
SCS_B200_FUSE_K34=1 REPS=20 timeout 300 python scripts/prof_spmv.py 2>&1 | tail -3 >> $L
echo "=== bench TTE" >> $L
SCS_BENCH_TTE=1 timeout 900 python bench.py --steps 100 --warmup 5 --no-cpu-baseline 2>gpurun_out/r02a_bench.err | tail -1 > gpurun_out/r02a_bench.json
cat gpurun_out/r02a_bench.json >> $L
echo "=== full gated suite" >> $L
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 >> $L
cat $L
