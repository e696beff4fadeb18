#!/bin/bash
# GPU call 1: microbenchmarks, ncu source profile of pipe4096, GPU tests, bench.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
( time timeout 300 ./tools/issuebench2 ) > gpurun_out/issuebench2.txt 2>&1
( time timeout 600 python tools/kbench.py --n 4096 2048 512 65536 --mb 2048 ) > gpurun_out/kbench.txt 2>&1
( time timeout 600 ncu --set full --clock-control none --import-source on -k regex:pipe4096 -s 1 -c 1 -f -o gpurun_out/prof_pipe4096 \
    python tools/kbench.py --n 4096 --mb 1024 --reps 1 --no-check ) > gpurun_out/ncu_full.log 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.txt 2>&1
( time timeout 600 python bench.py ) > gpurun_out/bench_n1.txt 2>&1
tail -3 gpurun_out/pytest_gpu.txt
tail -2 gpurun_out/bench_n1.txt
cat gpurun_out/kbench.txt
