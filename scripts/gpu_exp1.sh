#!/bin/bash
# experiment 1 (one GPU): gather ceiling, v3 parity, v3 vs v2 timing, v3 variants. Everything -> gpurun_out/exp1.log
mkdir -p gpurun_out
L=gpurun_out/exp1.log
: > $L
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv >> $L
echo "=== gather ceiling" >> $L
timeout 300 ./scripts/bin/gather_bench > gpurun_out/gather_bench.txt 2>&1; tail -3 gpurun_out/gather_bench.txt >> $L
echo "=== pytest linsys (v3 default)" >> $L
timeout 900 python -m pytest tests/test_linsys_gpu.py -x -q -m gpu 2>&1 | tail -15 >> $L
echo "=== v3 default (CHECK)" >> $L
CHECK=1 REPS=20 timeout 300 python scripts/prof_spmv.py 2>&1 | tail -5 >> $L
echo "=== v2" >> $L
SCS_B200_SPMV=2 REPS=20 timeout 300 python scripts/prof_spmv.py 2>&1 | tail -3 >> $L
for lib in scs_b200/variants/libscs_b200_*.so; do
  echo "== $lib" >> $L
  SCS_B200_LIB=$PWD/$lib REPS=20 timeout 300 python scripts/prof_spmv.py 2>&1 | tail -3 >> $L
done
cat $L
