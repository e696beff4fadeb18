#!/bin/bash
mkdir -p gpurun_out
{
echo "== small-bin tests"; timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -x -q -k "batched_small or all_sizes or kat_zero or straddling or golden or parseval or config1" 2>&1 | tail -6
echo "== batched 16..256"; timeout -s KILL 100 python tools/kbench.py --n 16 32 64 128 256 --mb 2048 --reps 5
echo "== batched 16..256 window"; timeout -s KILL 100 python tools/kbench.py --n 16 64 256 --mb 2048 --reps 5 --window
echo "== small_kernel 16 32"; timeout -s KILL 100 python tools/kbench.py --n 16 32 --mb 512 --reps 3 --flags 4
} > gpurun_out/step9.txt 2>&1
cat gpurun_out/step9.txt
