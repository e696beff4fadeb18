#!/bin/bash
# 8192 / 16384 bins: hand-off through tensor memory (no scratch, no helper warps) against the L2-scratch version
mkdir -p gpurun_out
{
for n in 8192 16384; do
  echo "== sanity $n"; timeout -s KILL 60 python tools/kbench.py --n $n --reps 3 --mb 256 --flags 64 || { echo "HANG/FAIL at $n: stop"; exit 1; }
done
echo "== solo tests"; timeout -s KILL 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "solo or mixes" 2>&1 | tail -6
for n in 8192 16384; do
  echo "== $n tensor-memory hand-off / L2 scratch / window"
  timeout -s KILL 60 python tools/kbench.py --n $n --reps 3 --mb 4096 --flags 64 --no-check
  timeout -s KILL 60 python tools/kbench.py --n $n --reps 3 --mb 4096 --flags 64 --no-check --lib variants/librpf_notm.so
  timeout -s KILL 60 python tools/kbench.py --n $n --reps 3 --mb 4096 --flags 64 --no-check --window
done
echo "== 4096 for comparison"; timeout -s KILL 60 python tools/kbench.py --n 4096 --reps 3 --mb 4096 --no-check
} > gpurun_out/step17.txt 2>&1
cat gpurun_out/step17.txt
