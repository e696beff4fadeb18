#!/bin/bash
# solo65536: L2 residency variants (eviction hints, persisting set-aside), CTA-count sweep, stall profile
mkdir -p gpurun_out
KB="timeout -s KILL 90 python tools/kbench.py --n 65536 --mb 4096 --reps 3 --flags 64"
{
echo "== solo tests"; timeout -s KILL 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "solo65536" 2>&1 | tail -5
echo "== solo (no hints, ring 2)"; $KB
for v in hints hintsp persist rawstg4; do echo "== solo $v"; $KB --lib variants/librpf_$v.so 2>&1 | tail -2; done
for g in 37 50 60 66; do echo "== solo hintsp, $((2*g)) CTAs"; $KB --scratch-mib $g --no-check --lib variants/librpf_hintsp.so 2>&1 | tail -1; done
for g in 37 60; do echo "== solo hints, $((2*g)) CTAs"; $KB --scratch-mib $g --no-check --lib variants/librpf_hints.so 2>&1 | tail -1; done
echo "== solo hints window"; $KB --window --lib variants/librpf_hints.so
echo "== default window"; timeout -s KILL 90 python tools/kbench.py --n 65536 --mb 4096 --reps 3 --window
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,lts__t_sector_hit_rate.pct
for cfg in "hintsp 0" "hintsp 37" "hints 0"; do set -- $cfg
  echo "== ncu traffic $1 scratch-mib $2"
  timeout -s KILL 200 ncu --metrics $M --clock-control none -k regex:solo65536 -c 1 --csv --log-file gpurun_out/solo_traffic_$1_$2.csv python tools/kbench.py --n 65536 --mb 2048 --reps 1 --flags 64 --no-check --scratch-mib $2 --lib variants/librpf_$1.so > /dev/null 2>&1
  grep -v "^==" gpurun_out/solo_traffic_$1_$2.csv | awk -F'","' '{print $(NF-2), $(NF)}' | tail -5
done
echo "== ncu full (hints, 148 CTAs)"
timeout -s KILL 300 ncu --set full --import-source on --clock-control none -k regex:solo65536 -c 1 -o gpurun_out/solo_full -f python tools/kbench.py --n 65536 --mb 1024 --reps 1 --flags 64 --no-check --lib variants/librpf_hints.so > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
} > gpurun_out/step11.txt 2>&1
cat gpurun_out/step11.txt
