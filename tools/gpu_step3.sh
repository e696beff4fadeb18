#!/bin/bash
# round-2 step: pipen (slot barriers) + ring depth A/B, TMA outer pass A/B, tests -- short timeouts
mkdir -p gpurun_out
{
echo "== pipen tests"; timeout -s KILL 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_host.py -m gpu -x -q -k "pipelined_small_bins or nondeterminism or tma_outer" 2>&1 | tail -8
for lib in "" "--lib variants/librpf_st3.so" "--lib variants/librpf_st4.so"; do
  echo "== pipeline $lib"; timeout -s KILL 60 python tools/kbench.py --n 512 1024 2048 --mb 4096 --reps 5 --flags 8 --no-check $lib
  echo "== pipeline window $lib"; timeout -s KILL 60 python tools/kbench.py --n 512 1024 2048 --mb 4096 --reps 5 --flags 8 --window --no-check $lib
done
echo "== four-step, TMA outer"; timeout -s KILL 90 python tools/kbench.py --n 8192 16384 32768 65536 --mb 4096 --reps 5
echo "== four-step, TMA outer, window"; timeout -s KILL 90 python tools/kbench.py --n 8192 65536 --mb 4096 --reps 5 --window
echo "== four-step, plain outer"; timeout -s KILL 90 python tools/kbench.py --n 8192 16384 32768 65536 --mb 4096 --reps 5 --flags 4 --no-check
echo "== launches 65536"; timeout -s KILL 120 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 12 --csv --log-file gpurun_out/r2_fourstep_launches.csv python tools/kbench.py --n 65536 --mb 256 --reps 1 --no-check > /dev/null 2>&1; tail -4 gpurun_out/r2_fourstep_launches.csv | cut -c1-400
echo "== tests"; timeout -s KILL 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
} > gpurun_out/step3.txt 2>&1
cat gpurun_out/step3.txt
