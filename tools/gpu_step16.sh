#!/bin/bash
# 65536 bins: layout 1 of the in-place scratch -- plain transpose [a][ko] (default) against the generic field exchange
mkdir -p gpurun_out
{
KB="timeout -s KILL 60 python tools/kbench.py --n 65536 --reps 3 --flags 64"
echo "== sanity"; $KB --mb 256 || { echo "HANG/FAIL: stop"; exit 1; }
echo "== transpose (default)"; $KB --mb 4096 --no-check
echo "== generic layout"; $KB --mb 4096 --no-check --lib variants/librpf_l1gen.so
echo "== transpose again"; $KB --mb 4096 --no-check
echo "== window"; $KB --mb 4096 --no-check --window
echo "== solo tests"; timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "solo or mixes" 2>&1 | tail -3
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,lts__t_sector_hit_rate.pct
for v in "" "--lib variants/librpf_l1gen.so"; do
echo "== ncu traffic $v"
timeout -s KILL 120 ncu --metrics $M --clock-control none -k regex:solo65536 -c 1 --csv --log-file gpurun_out/solo_traffic3.csv python tools/kbench.py --n 65536 --mb 2048 --reps 1 --flags 64 --no-check $v > /dev/null 2>&1
grep -v "^==" gpurun_out/solo_traffic3.csv | awk -F'","' '{print $(NF-2), $(NF)}' | tail -5
done
} > gpurun_out/step16.txt 2>&1
cat gpurun_out/step16.txt
