#!/bin/bash
for lib in scs_b200/variants/libscs_b200_*.so; do
  echo "== $lib"
  SCS_B200_LIB=$PWD/$lib REPS=20 timeout 300 python scripts/prof_spmv.py 2>&1 | tail -3
done
