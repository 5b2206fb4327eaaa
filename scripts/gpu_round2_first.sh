#!/bin/bash
# First GPU call of the next round (one GPU, ~10 min): (1) the tests staged as `gpu_unverified` at the end of round 1
# (reference fixtures through the device driver, exp/power goldens), (2) e2e / value with the threaded host setup,
# (3) the full gated suite. Everything -> gpurun_out/round2_first.log
mkdir -p gpurun_out
L=gpurun_out/round2_first.log
: > $L
echo "=== staged tests (-m gpu_unverified)" >> $L
timeout 600 python -m pytest tests -q -m gpu_unverified -s 2>&1 | tail -25 >> $L
echo "=== bench (no cpu baseline)" >> $L
SCS_BENCH_TTE=1 timeout 900 python bench.py --steps 100 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print({k:d[k] for k in ('value','ms_per_step','cg_iters_per_step','gpu_launches','lin_sys_ms','setup_ms')}, 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'tte', d.get('time_to_eps_1e-4'))" >> $L
echo "=== staged: K3+K4 fused (SCS_B200_FUSE_K34=1): linsys + solver tests, CG iteration timing" >> $L
SCS_B200_FUSE_K34=1 timeout 600 python -m pytest tests/test_linsys_gpu.py tests/test_solver_gpu.py -x -q -m gpu 2>&1 | tail -3 >> $L
REPS=20 timeout 300 python scripts/prof_spmv.py 2>&1 | tail -1 >> $L
SCS_B200_FUSE_K34=1 REPS=20 timeout 300 python scripts/prof_spmv.py 2>&1 | tail -1 >> $L
echo "=== full gated suite" >> $L
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 >> $L
cat $L
