#!/bin/bash
# round-2 step: cluster probe, cluster65536 + pipen checks, each under its own timeout
mkdir -p gpurun_out
{
echo "== probe"; timeout -s KILL 60 ./tools/cluster_probe
echo "== cluster tests"; timeout -s KILL 240 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cluster65536" 2>&1 | tail -15
echo "== kbench 65536"; timeout -s KILL 120 python tools/kbench.py --n 65536 --mb 4096 --reps 5
timeout -s KILL 120 python tools/kbench.py --n 65536 --mb 4096 --reps 5 --flags 16
echo "== pipen tests"; timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_host.py -m gpu -x -q -k "pipelined_small_bins or nondeterminism" 2>&1 | tail -15
echo "== kbench small"; timeout -s KILL 120 python tools/kbench.py --n 512 1024 2048 --mb 4096 --reps 5
timeout -s KILL 120 python tools/kbench.py --n 512 --mb 4096 --reps 5 --flags 8
} > gpurun_out/step1.txt 2>&1
cat gpurun_out/step1.txt
