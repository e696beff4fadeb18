#!/bin/bash
mkdir -p gpurun_out
{
echo "== cluster tests"; timeout -s KILL 120 python -m pytest tests/test_gpu_parity.py tests/test_gpu_host.py -m gpu -x -q -k "cluster65536 or (nondeterminism and 65536)" 2>&1 | tail -4
echo "== kbench 65536 cluster"; timeout -s KILL 60 python tools/kbench.py --n 65536 --mb 4096 --reps 5 --flags 16
echo "== kbench 65536 default"; timeout -s KILL 60 python tools/kbench.py --n 65536 --mb 4096 --reps 5
echo "== 4096 default"; timeout -s KILL 60 python tools/kbench.py --n 4096 --mb 4096 --reps 7
echo "== 4096 acc f64"; timeout -s KILL 60 python tools/kbench.py --n 4096 --mb 4096 --reps 7 --lib variants/librpf_accf64.so
echo "== 4096 acc f64 window"; timeout -s KILL 60 python tools/kbench.py --n 4096 --mb 4096 --reps 7 --window --lib variants/librpf_accf64.so
echo "== 65536 acc f64"; timeout -s KILL 60 python tools/kbench.py --n 65536 --mb 4096 --reps 5 --lib variants/librpf_accf64.so
} > gpurun_out/step6.txt 2>&1
cat gpurun_out/step6.txt
