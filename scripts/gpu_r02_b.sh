#!/bin/bash
# round 2, call B (1 GPU): new parity tests + full suite, PDL A/B, bench in the driver's form, ncu launch list + full capture of K1/K2
mkdir -p gpurun_out
L=gpurun_out/r02b.log
: > $L
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader >> $L
echo "=== CG kernels, PDL on / off" >> $L
timeout 300 python scripts/prof_cg.py >> $L 2>&1
SCS_B200_PDL=0 timeout 300 python scripts/prof_cg.py >> $L 2>&1
echo "=== new parity tests" >> $L
timeout 1500 python -m pytest tests/test_golden_gpu.py tests/test_parity_configs_gpu.py -q -s -m gpu 2>&1 | grep -v "^$" | tail -70 >> $L
echo "=== full gated suite" >> $L
timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 >> $L
echo "=== smoke" >> $L
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> $L 2>&1
echo "=== bench (driver form)" >> $L
timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 2>gpurun_out/r02b_bench.err | tail -1 > gpurun_out/r02b_bench.json
cat gpurun_out/r02b_bench.json >> $L
echo "=== ncu: launch list of a short solve" >> $L
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 1500 --csv \
    --log-file gpurun_out/r02b_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-tte > gpurun_out/r02b_ncu_bench.log 2>&1
echo "=== ncu: full capture of the in-loop SpMV kernels" >> $L
REPS=6 timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmv_flag_kernel -s 12 -c 4 \
    -o gpurun_out/r02b_spmv_inloop -f python scripts/prof_cg.py > gpurun_out/r02b_ncu_spmv.log 2>&1
tail -3 gpurun_out/r02b_ncu_spmv.log >> $L
cat $L
