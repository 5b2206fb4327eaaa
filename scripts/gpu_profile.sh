#!/bin/bash
# ncu evidence for the hot kernels (run under gpurun, ONE GPU). Outputs -> gpurun_out/
# Uses a reduced iteration count: ncu serialises and replays kernels.
mkdir -p gpurun_out
CFG=${CFG:-C2}
# 1) launch list with device time per launch (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 1500 --csv \
    --log-file gpurun_out/launches_${CFG}.csv \
    python bench.py --config $CFG --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_${CFG}.log 2>&1
# 2) full section set for the SpMV kernels (3 launches each, after warm-up)
ncu --set full --clock-control none --import-source on -k regex:spmv_csr_stream -s 200 -c 4 \
    -o gpurun_out/spmv_${CFG} -f \
    python bench.py --config $CFG --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_spmv_${CFG}.log 2>&1
# 3) the CG vector kernels
ncu --set full --clock-control none --import-source on -k regex:k_cg_ -s 200 -c 4 \
    -o gpurun_out/cgvec_${CFG} -f \
    python bench.py --config $CFG --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_cgvec_${CFG}.log 2>&1
ls -la gpurun_out
