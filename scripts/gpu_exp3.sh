#!/bin/bash
# experiment 3: v3 with coalesced epilogue: parity, timing vs v2, mixed mode, variants
mkdir -p gpurun_out
L=gpurun_out/exp3.log
: > $L
echo "=== pytest linsys (v3 default)" >> $L
timeout 900 python -m pytest tests/test_linsys_gpu.py -x -q -m gpu 2>&1 | tail -6 >> $L
echo "=== v3 default (CHECK)" >> $L
CHECK=1 REPS=20 timeout 300 python scripts/prof_spmv.py 2>&1 | tail -4 >> $L
echo "=== mixed: v2 for avg row < 6" >> $L
SCS_B200_SPMV_MINAVG=6 REPS=20 timeout 300 python scripts/prof_spmv.py 2>&1 | tail -3 >> $L
echo "=== v2" >> $L
SCS_B200_SPMV=2 REPS=20 timeout 300 python scripts/prof_spmv.py 2>&1 | tail -3 >> $L
for lib in scs_b200/variants/libscs_b200_*.so; do
  echo "== $lib" >> $L
  SCS_B200_LIB=$PWD/$lib REPS=20 timeout 300 python scripts/prof_spmv.py 2>&1 | tail -3 >> $L
done
cat $L
