#!/bin/bash
# round 2, call L (1 GPU): final validation -- the driver's three commands (gated suite, smoke, bench) on the final tree,
# then one bench line per other BASELINE config (no CPU baseline, no time-to-eps)
mkdir -p gpurun_out
L=gpurun_out/r02l.log
: > $L
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader >> $L
echo "=== full gated suite (the driver's command), timed" >> $L
SECONDS=0
timeout 2400 python -m pytest tests -x -q -m gpu --durations=8 2>&1 | tail -22 >> $L
echo "suite wall ${SECONDS} s" >> $L
echo "=== smoke" >> $L
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> $L 2>&1
echo "=== bench (driver form)" >> $L
timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 2>gpurun_out/r02l_bench.err | tail -1 > gpurun_out/r02l_bench.json
cat gpurun_out/r02l_bench.json >> $L
for cfg in C3 C4 C5; do
  echo "=== bench $cfg" >> $L
  timeout 900 python bench.py --config $cfg --steps 20 --warmup 5 --no-cpu-baseline --no-tte 2>/dev/null | tail -1 > gpurun_out/r02l_bench_$cfg.json
  python - >> $L <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02l_bench_$cfg.json").read())
    print("$cfg", {k: d[k] for k in ("value", "ms_per_step", "cg_iters_per_step", "lin_sys_ms", "cone_ms", "accel_ms", "setup_ms")}, "e2e", d["e2e"]["value"],
          [(r["kernel"][:14], round(r["ms"] * 1e3, 1), round(r["frac"], 3)) for r in d.get("roofline_all", [])])
except Exception as e:
    print("$cfg failed", e)
PY
done
cat $L
