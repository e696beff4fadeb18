#!/bin/bash
# last check of the round: smoke + golden vectors (incl. the one-read fixture that runs on the solo kernel) + the solo tests
mkdir -p gpurun_out
{
timeout -s KILL 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout -s KILL 100 python -m pytest tests/test_golden.py tests/test_gpu_parity.py -m gpu -q -k "golden or solo or mixes or page_locked" 2>&1 | tail -3
} > gpurun_out/step21.txt 2>&1
cat gpurun_out/step21.txt
