#!/bin/bash
# full GPU test suite (run under gpurun)
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -s -m gpu "$@" 2>&1 | tail -80
