#!/bin/bash
# round 2, call P (re-run of D on the final tree: allocator caches, device long rows) (N GPUs, default 2): sharded solve vs reference, CG iteration in the replicated modes and in the push-based
# sharded-x mode, one bench line per mode.   gpurun --gpus 2 -- bash scripts/gpu_r02_d.sh 2
N=${1:-2}   # MODES="1" (or "") limits the bench part: one line per SCS_B200_SHARD_X value
mkdir -p gpurun_out
L=gpurun_out/r02p_${N}gpu.log
: > $L
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8 >> $L
MGPU_SHARD_X=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    tests/mgpu_check.py 2>&1 | grep -v "^W\|warn\|^\*\|OMP_NUM" | tail -14 >> $L
for mode in ${MODES:-1 0}; do
  SCS_B200_SHARD_X=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$mode \
      bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-tte > gpurun_out/r02p_bench_${N}gpu_shardx$mode.json 2> gpurun_out/r02p_bench_${N}gpu_shardx$mode.err
  grep -v "^\*\|OMP_NUM\|^$\|^W" gpurun_out/r02p_bench_${N}gpu_shardx$mode.err | tail -4 >> $L
  python - >> $L <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/r02p_bench_${N}gpu_shardx$mode.json") if l.startswith("{")][-1])
    print("SHARD_X=$mode", {k: d[k] for k in ("n_gpus", "value", "ms_per_step", "cg_iters_per_step", "setup_ms")}, "e2e", d["e2e"]["value"])
except Exception as e:
    print("no bench json", e)
PY
done
cat $L
