#!/bin/bash
mkdir -p gpurun_out
{
echo "== ws"; timeout 100 python tools/kbench.py --n 65536 --mb 1024 --reps 3; echo "rc=$?"
timeout 100 python tools/kbench.py --n 65536 8192 16384 32768 --mb 4096 --reps 5; echo "rc=$?"
timeout 100 python tools/kbench.py --n 65536 --mb 4096 --reps 5 --window; echo "rc=$?"
echo "== split"; RPF_B200_FOURSTEP=split timeout 100 python tools/kbench.py --n 65536 --mb 4096 --reps 5; echo "rc=$?"
if [ "$1" = "test" ]; then timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5; fi
nvidia-smi --query-gpu=name,memory.used --format=csv
} > gpurun_out/fs.txt 2>&1
cat gpurun_out/fs.txt
