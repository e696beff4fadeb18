#!/bin/bash
mkdir -p gpurun_out
{
timeout 120 python tools/kbench.py --n 65536 8192 16384 32768 --mb 4096 --reps 5; echo "rc=$?"
timeout 120 python tools/kbench.py --n 65536 --mb 4096 --reps 5 --window; echo "rc=$?"
for v in variants/*.so; do
  [ -f "$v" ] || continue
  echo "== $v"; RPF_B200_LIB=$PWD/$v timeout 120 python tools/kbench.py --n 65536 --mb 4096 --reps 5; echo "rc=$?"
done
if [ "$1" = "test" ]; then timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5; fi
} > gpurun_out/fs.txt 2>&1
cat gpurun_out/fs.txt
