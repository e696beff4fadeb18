#!/bin/bash
# A/B: default library vs variants/*.so on the kernel micro-benchmark (+ optional GPU parity tests).
mkdir -p gpurun_out
{
echo "== default"; timeout -s KILL 200 python tools/kbench.py --n 4096 --mb 4096 --reps 7
for v in variants/*.so; do
  echo "== $v"; timeout -s KILL 60 python tools/kbench.py --n 4096 --mb 4096 --reps 7 --lib $v
done
if [ "$1" = "test" ]; then timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5; fi
} > gpurun_out/ab.txt 2>&1
cat gpurun_out/ab.txt
