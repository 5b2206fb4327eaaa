#!/bin/bash
# round 2, call Q (1 GPU, last minutes of the budget): smoke + the test files not re-run since the last refactor
mkdir -p gpurun_out
L=gpurun_out/r02q.log
: > $L
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" >> $L 2>&1
timeout 200 python -m pytest tests/test_linsys_gpu.py tests/test_golden_gpu.py tests/test_aa_gpu.py tests/test_reference_suite_gpu.py tests/test_zz_reference_fixtures_gpu.py -x -q -m gpu 2>&1 | tail -5 >> $L
cat $L
