#!/bin/bash
# one bench line per BASELINE config (single GPU), no CPU baseline (the reference needs minutes per iteration here)
mkdir -p gpurun_out
for CFG in "$@"; do
  echo "=== $CFG"
  timeout 1200 python bench.py --config $CFG --steps ${STEPS:-30} --warmup 5 --no-cpu-baseline \
      > gpurun_out/bench_${CFG}.json 2> gpurun_out/bench_${CFG}.err
  tail -2 gpurun_out/bench_${CFG}.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_${CFG}.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "cg_iters_per_step", "lin_sys_ms", "cone_ms", "accel_ms", "setup_ms")}, d["e2e"]["value"], d["config"]["n"], d["config"]["m"], d["config"]["nnz"])
    print([ (r["kernel"][:40], round(r["ms"],4), round(r["frac"],3)) for r in d.get("roofline_all", [])])
except Exception as e:
    print("no json", e)
PY
done
