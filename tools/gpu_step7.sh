#!/bin/bash
mkdir -p gpurun_out
{
echo "== batched tests"; timeout -s KILL 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_host.py -m gpu -x -q -k "batched_small or (nondeterminism and 32)" 2>&1 | tail -6
echo "== batched"; timeout -s KILL 60 python tools/kbench.py --n 512 1024 2048 --mb 4096 --reps 5 --flags 32
echo "== batched window"; timeout -s KILL 60 python tools/kbench.py --n 512 1024 2048 --mb 4096 --reps 5 --flags 32 --window
echo "== default"; timeout -s KILL 60 python tools/kbench.py --n 512 1024 2048 4096 --mb 4096 --reps 5 --no-check
echo "== default window"; timeout -s KILL 60 python tools/kbench.py --n 512 1024 2048 4096 --mb 4096 --reps 5 --no-check --window
} > gpurun_out/step7.txt 2>&1
cat gpurun_out/step7.txt
