#!/bin/bash
# scaling bench only (run with: gpurun --gpus N -- bash scripts/gpu_mgpu_bench.sh N [steps])
N=${1:-2}
STEPS=${2:-40}
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps $STEPS --warmup 5 --no-cpu-baseline > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/bench_${N}gpu.err | tail -4
python - <<PY
import json
d = json.load(open("gpurun_out/bench_${N}gpu.json"))
print({k: d[k] for k in ("n_gpus", "value", "ms_per_step", "cg_iters_per_step", "setup_ms")}, "e2e", d["e2e"]["value"])
PY
