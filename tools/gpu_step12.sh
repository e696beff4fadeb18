#!/bin/bash
# solo65536 with rows staged by the helper warps and interleaved stores: parity, speed, profile.  Stops at the first hang.
mkdir -p gpurun_out
KB="timeout -s KILL 60 python tools/kbench.py --n 65536 --reps 3 --flags 64"
{
echo "== sanity"; $KB --mb 256 || { echo "HANG/FAIL: stop"; exit 1; }
echo "== solo tests"; timeout -s KILL 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "solo65536" 2>&1 | tail -5
echo "== solo"; $KB --mb 4096 || exit 1
for g in 37; do echo "== solo, $((2*g)) CTAs"; $KB --mb 4096 --scratch-mib $g --no-check 2>&1 | tail -1; done
echo "== solo window"; $KB --mb 4096 --window
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,lts__t_sector_hit_rate.pct
echo "== ncu traffic"
timeout -s KILL 120 ncu --metrics $M --clock-control none -k regex:solo65536 -c 1 --csv --log-file gpurun_out/solo_traffic2.csv python tools/kbench.py --n 65536 --mb 2048 --reps 1 --flags 64 --no-check > /dev/null 2>&1
grep -v "^==" gpurun_out/solo_traffic2.csv | awk -F'","' '{print $(NF-2), $(NF)}' | tail -5
echo "== ncu full"
timeout -s KILL 200 ncu --set full --import-source on --clock-control none -k regex:solo65536 -c 1 -o gpurun_out/solo_full2 -f python tools/kbench.py --n 65536 --mb 1024 --reps 1 --flags 64 --no-check > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
} > gpurun_out/step12.txt 2>&1
cat gpurun_out/step12.txt
