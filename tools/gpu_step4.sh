#!/bin/bash
mkdir -p gpurun_out
{
for mib in 36 72 144 288 576; do
  echo "== 65536, scratch $mib MiB"; timeout -s KILL 60 python tools/kbench.py --n 65536 --mb 4608 --reps 5 --no-check --scratch-mib $mib
done
echo "== 8192, scratch 72 / 144"; timeout -s KILL 60 python tools/kbench.py --n 8192 --mb 4608 --reps 5 --no-check --scratch-mib 72
timeout -s KILL 60 python tools/kbench.py --n 8192 --mb 4608 --reps 5 --no-check --scratch-mib 144
echo "== small"; timeout -s KILL 60 python tools/kbench.py --n 512 1024 2048 --mb 4096 --reps 5
} > gpurun_out/step4.txt 2>&1
cat gpurun_out/step4.txt
