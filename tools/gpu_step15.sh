#!/bin/bash
# solo kernel generalised over the outer radix (8192 ... 65536 bins): parity + speed against the two-kernel path
mkdir -p gpurun_out
{
for n in 65536 32768 16384 8192; do
  echo "== sanity $n"; timeout -s KILL 60 python tools/kbench.py --n $n --reps 3 --mb 256 --flags 64 || { echo "HANG/FAIL at $n: stop"; exit 1; }
done
echo "== solo tests"; timeout -s KILL 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "solo" 2>&1 | tail -6
for n in 65536 32768 16384 8192; do
  echo "== $n solo / two-kernel / solo window"
  timeout -s KILL 60 python tools/kbench.py --n $n --reps 3 --mb 4096 --flags 64 --no-check
  timeout -s KILL 60 python tools/kbench.py --n $n --reps 3 --mb 4096 --flags 128 --no-check
  timeout -s KILL 60 python tools/kbench.py --n $n --reps 3 --mb 4096 --flags 64 --no-check --window
done
} > gpurun_out/step15.txt 2>&1
cat gpurun_out/step15.txt
