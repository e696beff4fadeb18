#!/bin/bash
# round-2 step: pipen ring depth A/B (1024 / 2048, with and without a window), then the whole GPU suite
mkdir -p gpurun_out
{
for lib in "" "--lib variants/librpf_st3.so" "--lib variants/librpf_st4.so"; do
  echo "== pipeline $lib"; timeout -s KILL 120 python tools/kbench.py --n 512 1024 2048 --mb 4096 --reps 5 --flags 8 --no-check $lib
  echo "== pipeline window $lib"; timeout -s KILL 120 python tools/kbench.py --n 512 1024 2048 --mb 4096 --reps 5 --flags 8 --window --no-check $lib
done
echo "== lockstep"; timeout -s KILL 120 python tools/kbench.py --n 512 1024 2048 --mb 4096 --reps 5 --flags 4 --no-check
echo "== lockstep window"; timeout -s KILL 120 python tools/kbench.py --n 512 1024 2048 --mb 4096 --reps 5 --flags 4 --window --no-check
echo "== tests"; timeout -s KILL 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
} > gpurun_out/step2.txt 2>&1
cat gpurun_out/step2.txt
