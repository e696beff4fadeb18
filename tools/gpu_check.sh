#!/bin/bash
# kernel micro-benchmark over the main sizes (+ optional GPU tests)
mkdir -p gpurun_out
{
timeout -s KILL 200 python tools/kbench.py --n 4096 512 2048 8192 65536 --mb 4096 --reps 5
timeout -s KILL 100 python tools/kbench.py --n 4096 65536 --mb 4096 --reps 5 --window
if [ "$1" = "test" ]; then timeout -s KILL 400 python -m pytest tests -m gpu -q 2>&1 | tail -3; fi
} > gpurun_out/check.txt 2>&1
cat gpurun_out/check.txt
