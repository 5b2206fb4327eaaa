#!/bin/bash
# round 2, call N (1 GPU): device builder with long rows, cuSOLVER handle cache, pinned-block cache; where scs_init /
# scs_finish spend their time (SCS_B200_SETUP_TIMING marks) on C2 / C4 / C3
mkdir -p gpurun_out
L=gpurun_out/r02n.log
: > $L
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader >> $L
echo "=== tests touching the changed code" >> $L
timeout 900 python -m pytest tests/test_linsys_gpu.py tests/test_cones_gpu.py tests/test_solver_gpu.py -x -q -m gpu 2>&1 | tail -6 >> $L
timeout 600 python -m pytest tests/test_parity_configs_gpu.py -x -q -m gpu -k "one_iteration or psd_c4" 2>&1 | tail -4 >> $L
for cfg in C2 C4 C5 C3; do
  echo "=== bench $cfg" >> $L
  timeout 900 python bench.py --config $cfg --steps 20 --warmup 5 --no-cpu-baseline --no-tte 2>/dev/null | tail -1 > gpurun_out/r02n_bench_$cfg.json
  python - >> $L <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02n_bench_$cfg.json").read())
    print("$cfg", {k: d[k] for k in ("value", "ms_per_step", "cg_iters_per_step", "setup_ms")}, "e2e", d["e2e"]["value"], d["e2e"].get("breakdown"),
          [(r["kernel"][:14], round(r["ms"] * 1e3, 1), round(r["frac"], 3)) for r in d.get("roofline_all", [])])
except Exception as e:
    print("$cfg failed", e)
PY
  echo "--- $cfg with SCS_B200_SETUP_TIMING=1 (every mark synchronises; last scs() call = the timed e2e one)" >> $L
  SCS_B200_SETUP_TIMING=1 timeout 900 python bench.py --config $cfg --steps 3 --warmup 3 --no-cpu-baseline --no-tte 2>&1 >/dev/null | grep "scs_b200 setup" | tail -64 >> $L
done
cat $L
