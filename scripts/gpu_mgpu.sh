#!/bin/bash
# multi-GPU: row-sharded solve check + scaling bench (run with: gpurun --gpus N -- bash scripts/gpu_mgpu.sh N)
N=${1:-2}
STEPS=${STEPS:-50}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    tests/mgpu_check.py 2>&1 | grep -v "^W\|warn" | tail -12
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps $STEPS --warmup 5 --no-cpu-baseline > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
tail -3 gpurun_out/bench_${N}gpu.err
cat gpurun_out/bench_${N}gpu.json
