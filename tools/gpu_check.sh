#!/bin/bash
# kernel micro-benchmark over the main sizes + GPU tests + bench
mkdir -p gpurun_out
{
timeout 600 python tools/kbench.py --n 4096 512 1024 2048 8192 65536 --mb 4096 --reps 5
timeout 300 python tools/kbench.py --n 4096 65536 --mb 4096 --reps 5 --window
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
} > gpurun_out/check.txt 2>&1
cat gpurun_out/check.txt
