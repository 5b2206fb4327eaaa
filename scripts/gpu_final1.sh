#!/bin/bash
# round-1 validation: full GPU test suite, smoke, v3 mini-sweep, ncu capture of the SpMV, bench (both arms)
mkdir -p gpurun_out
L=gpurun_out/final1.log
: > $L
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader >> $L; nproc >> $L
echo "=== pytest -m gpu" >> $L
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 >> $L
echo "=== smoke" >> $L
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 >> $L
echo "=== default" >> $L
REPS=20 timeout 300 python scripts/prof_spmv.py 2>&1 | tail -3 >> $L
for lib in scs_b200/variants/libscs_b200_*.so; do
  echo "== $lib" >> $L
  SCS_B200_LIB=$PWD/$lib REPS=20 timeout 300 python scripts/prof_spmv.py 2>&1 | tail -3 >> $L
done
echo "=== ncu" >> $L
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmv_ -s 8 -c 4 \
    -o gpurun_out/spmv_r01f -f python scripts/prof_spmv.py > gpurun_out/ncu_spmv_r01f.log 2>&1
tail -2 gpurun_out/ncu_spmv_r01f.log >> $L
echo "=== bench ours" >> $L
timeout 900 python bench.py --steps 100 --warmup 5 > gpurun_out/bench_ours_r01f.json 2> gpurun_out/bench_ours_r01f.err
tail -2 gpurun_out/bench_ours_r01f.err >> $L
cat gpurun_out/bench_ours_r01f.json >> $L
echo "=== bench reference" >> $L
timeout 600 python bench.py --impl reference --steps 100 --warmup 5 > gpurun_out/bench_ref_r01f.json 2> gpurun_out/bench_ref_r01f.err
cat gpurun_out/bench_ref_r01f.json >> $L
cat $L | cut -c1-1500
