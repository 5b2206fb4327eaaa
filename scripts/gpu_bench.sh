#!/bin/bash
# bench + smoke on the GPU box; outputs go to gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
nproc; lscpu | grep "Model name"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
STEPS=${STEPS:-100}
timeout 1500 python bench.py --steps $STEPS --warmup 5 "$@" > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
tail -5 gpurun_out/bench_ours.err
cat gpurun_out/bench_ours.json
