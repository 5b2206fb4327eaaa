#!/bin/bash
# round 2, call M (1 GPU): validation of the final tree (caching device allocator, C3 tests) -- the driver's three
# commands, then the e2e line with the allocator cache off for comparison
mkdir -p gpurun_out
L=gpurun_out/r02m.log
: > $L
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader >> $L
echo "=== full gated suite (the driver's command), timed" >> $L
SECONDS=0
timeout 2400 python -m pytest tests -x -q -m gpu --durations=8 2>&1 | tail -22 >> $L
echo "suite wall ${SECONDS} s" >> $L
echo "=== C3 tests, verbose" >> $L
timeout 600 python -m pytest tests/test_parity_configs_gpu.py -q -s -m gpu -k "c3" 2>&1 | grep -v "^ERROR" | tail -12 >> $L
echo "=== smoke" >> $L
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> $L 2>&1
echo "=== bench (driver form)" >> $L
timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 2>gpurun_out/r02m_bench.err | tail -1 > gpurun_out/r02m_bench.json
cat gpurun_out/r02m_bench.json >> $L
echo "=== bench, allocator cache off" >> $L
SCS_B200_POOL=0 timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-tte 2>/dev/null | tail -1 > gpurun_out/r02m_bench_nopool.json
python - >> $L <<PY
import json
for f in ("r02m_bench", "r02m_bench_nopool"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read())
        print(f, "value", d["value"], "e2e", d["e2e"])
    except Exception as e:
        print(f, "failed", e)
PY
cat $L
