#!/bin/bash
# round 2, call C (1 GPU): device-side setup (timing, host fallback), reordered CG pair on/off, tests of the new paths, bench
mkdir -p gpurun_out
L=gpurun_out/r02c.log
: > $L
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader >> $L
echo "=== CG kernels: reorder on (default) / off; PDL on" >> $L
timeout 300 python scripts/prof_cg.py >> $L 2>&1
SCS_B200_REORDER=0 timeout 300 python scripts/prof_cg.py >> $L 2>&1
echo "=== setup timing: device path, then host path" >> $L
SCS_B200_SETUP_TIMING=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-tte 2>&1 | grep "scs_b200 setup" | tail -12 >> $L
echo "--- host path" >> $L
SCS_B200_HOST_SETUP=1 SCS_B200_SETUP_TIMING=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-tte 2>&1 | grep "scs_b200 setup" | tail -12 >> $L
echo "=== tests of the new paths" >> $L
timeout 1500 python -m pytest tests/test_linsys_gpu.py tests/test_solver_gpu.py tests/test_golden_gpu.py tests/test_cones_gpu.py -x -q -m gpu 2>&1 | tail -8 >> $L
echo "--- host setup path, no reorder" >> $L
SCS_B200_HOST_SETUP=1 SCS_B200_REORDER=0 timeout 900 python -m pytest tests/test_linsys_gpu.py tests/test_golden_gpu.py -x -q -m gpu 2>&1 | tail -4 >> $L
echo "=== parity configs" >> $L
timeout 1500 python -m pytest tests/test_parity_configs_gpu.py -q -s -m gpu 2>&1 | grep -E "^\[|passed|failed|Error|error" | tail -30 >> $L
echo "=== bench (driver form)" >> $L
timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 2>gpurun_out/r02c_bench.err | tail -1 > gpurun_out/r02c_bench.json
cat gpurun_out/r02c_bench.json >> $L
cat $L
