#!/bin/bash
# round 2, call E (1 GPU): K1 with the reciprocal multiply, setup breakdown, FULL gated suite (timed), bench in the driver's form,
# C4 bench with the DMMA / FMA contraction, ncu of the DMMA kernel (tensor pipe)
mkdir -p gpurun_out
L=gpurun_out/r02e.log
: > $L
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader >> $L
echo "=== CG kernels (K1 = POST_MUL)" >> $L
timeout 300 python scripts/prof_cg.py >> $L 2>&1
echo "=== setup timing: device path" >> $L
SCS_B200_SETUP_TIMING=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-tte 2>&1 | grep "scs_b200 setup" | tail -14 >> $L
echo "=== full gated suite (timed)" >> $L
/usr/bin/time -f "suite wall %e s" timeout 2400 python -m pytest tests -x -q -m gpu --durations=12 2>&1 | tail -30 >> $L
echo "=== smoke" >> $L
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> $L 2>&1
echo "=== bench (driver form)" >> $L
timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 2>gpurun_out/r02e_bench.err | tail -1 > gpurun_out/r02e_bench.json
cat gpurun_out/r02e_bench.json >> $L
echo "=== C4 (SDP 200 x PSD(100)): DMMA vs FMA contraction" >> $L
for v in 0 1; do
  SCS_B200_PSD_FMA=$v timeout 900 python bench.py --config C4 --steps 40 --warmup 5 --no-cpu-baseline --no-tte 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('PSD_FMA=$v', {k:d[k] for k in ('value','ms_per_step','lin_sys_ms','cone_ms','accel_ms','cg_iters_per_step')})" >> $L
done
echo "=== ncu: tensor pipe of the DMMA contraction" >> $L
timeout 900 ncu --metrics sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,gpu__time_duration.sum \
    --clock-control none -k regex:k_psd_reconstruct -c 2 --csv --log-file gpurun_out/r02e_ncu_dmma.csv \
    python -m pytest tests/test_parity_configs_gpu.py -q -m gpu -k psd_c4 > gpurun_out/r02e_ncu_dmma.log 2>&1
grep -E "k_psd_reconstruct" gpurun_out/r02e_ncu_dmma.csv | cut -c1-400 | tail -6 >> $L
cat $L
