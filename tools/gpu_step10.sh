#!/bin/bash
# solo65536 (one CTA per spectrum, thread-private hand-off): first run -- parity, speed against the two-kernel path, DRAM traffic
mkdir -p gpurun_out
{
echo "== solo tests"; timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "solo65536" 2>&1 | tail -15
echo "== kbench 65536 solo"; timeout -s KILL 90 python tools/kbench.py --n 65536 --mb 4096 --reps 5 --flags 64
echo "== kbench 65536 default"; timeout -s KILL 90 python tools/kbench.py --n 65536 --mb 4096 --reps 5
for g in 37 50 60 66; do echo "== kbench 65536 solo, scratch $g MiB = $((2*g)) CTAs"; timeout -s KILL 90 python tools/kbench.py --n 65536 --mb 4096 --reps 3 --flags 64 --scratch-mib $g --no-check; done
for v in rawstg2 flush32 hints; do
  [ -f variants/librpf_$v.so ] && { echo "== kbench 65536 solo $v"; timeout -s KILL 90 python tools/kbench.py --n 65536 --mb 4096 --reps 5 --flags 64 --lib variants/librpf_$v.so; }
done
echo "== ncu traffic solo"
timeout -s KILL 240 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:solo65536 -c 2 --csv --log-file gpurun_out/solo_traffic.csv python tools/kbench.py --n 65536 --mb 2048 --reps 1 --flags 64 --no-check > /dev/null 2>&1
grep -v "^==" gpurun_out/solo_traffic.csv | cut -d, -f5,13- | tail -12
} > gpurun_out/step10.txt 2>&1
cat gpurun_out/step10.txt
