#!/bin/bash
# round 2, call F (1 GPU): the FULL gated suite, timed (driver's command), incl. the DMMA TSQR of Anderson acceleration; AA A/B
mkdir -p gpurun_out
L=gpurun_out/r02f.log
: > $L
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader >> $L
echo "=== AA parity tests: DMMA TSQR (default), then FMA TSQR" >> $L
timeout 600 python -m pytest tests/test_aa_gpu.py tests/test_golden_gpu.py -q -s -m gpu -k "aa or anderson" 2>&1 | grep -E "^\[|passed|failed|rror" | tail -12 >> $L
SCS_B200_AA_FMA=1 timeout 600 python -m pytest tests/test_aa_gpu.py -q -m gpu 2>&1 | tail -2 >> $L
echo "=== full gated suite (the driver's command), timed" >> $L
SECONDS=0
timeout 2400 python -m pytest tests -x -q -m gpu --durations=15 2>&1 | tail -32 >> $L
echo "suite wall ${SECONDS} s" >> $L
echo "=== smoke" >> $L
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> $L 2>&1
cat $L
