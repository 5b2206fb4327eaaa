#!/bin/bash
# first GPU smoke: linsys parity + a quick SpMV/CG timing
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_linsys_gpu.py -x -q -s -m gpu 2>&1 | tail -40
