#!/bin/bash
mkdir -p gpurun_out
{
echo "== 256-bin tests"; timeout -s KILL 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "batched_small or all_sizes or kat_zero or straddling or golden" 2>&1 | tail -4
echo "== 256 batched (default)"; timeout -s KILL 60 python tools/kbench.py --n 256 --mb 4096 --reps 5
echo "== 256 batched window"; timeout -s KILL 60 python tools/kbench.py --n 256 --mb 4096 --reps 5 --window
echo "== 256/128/64 small_kernel"; timeout -s KILL 100 python tools/kbench.py --n 256 --mb 1024 --reps 3 --flags 4
timeout -s KILL 100 python tools/kbench.py --n 128 64 --mb 1024 --reps 3
} > gpurun_out/step8.txt 2>&1
cat gpurun_out/step8.txt
