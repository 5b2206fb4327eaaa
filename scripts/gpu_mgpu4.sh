#!/bin/bash
# multi-GPU validation (run with: gpurun --gpus N -- bash scripts/gpu_mgpu4.sh N): sharded solve vs reference,
# CG-iteration timing in both peer-memory reduction modes, one bench line
N=${1:-4}
mkdir -p gpurun_out
L=gpurun_out/mgpu${N}.log
: > $L
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8 >> $L
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    tests/mgpu_check.py 2>&1 | grep -v "^W\|warn\|^\*\|OMP_NUM" | tail -12 >> $L
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps ${STEPS:-40} --warmup 5 --no-cpu-baseline > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
grep -v "^\*\|OMP_NUM\|^$\|^W" gpurun_out/bench_${N}gpu.err | tail -4 >> $L
python - >> $L <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/bench_${N}gpu.json") if l.startswith("{")][-1])
    print({k: d[k] for k in ("n_gpus", "value", "ms_per_step", "cg_iters_per_step", "setup_ms")}, "e2e", d["e2e"]["value"])
    print([(r["kernel"][:30], round(r["ms"] * 1e3, 1)) for r in d.get("roofline_all", [])])
except Exception as e:
    print("no bench json", e)
PY
cat $L
