#!/bin/bash
mkdir -p gpurun_out
{
for mib in 1024 2048 4096; do
  echo "== 65536, scratch $mib MiB"; timeout -s KILL 60 python tools/kbench.py --n 65536 --mb 8192 --reps 5 --no-check --scratch-mib $mib
done
for mib in 256 1024 4096; do
echo "== 8192/16384/32768, scratch $mib"; timeout -s KILL 90 python tools/kbench.py --n 8192 16384 32768 --mb 8192 --reps 5 --no-check --scratch-mib $mib
done
echo "== tests"; timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fourstep or tma_outer or 65536 or all_sizes or straddling" 2>&1 | tail -4
} > gpurun_out/step5.txt 2>&1
cat gpurun_out/step5.txt
