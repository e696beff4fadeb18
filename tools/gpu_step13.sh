#!/bin/bash
# validation of the new 65536-bin default: whole GPU suite, bench line, launch list with DRAM bytes, sanitizers on the solo kernel
mkdir -p gpurun_out
{
echo "== sanity"; timeout -s KILL 60 python tools/kbench.py --n 65536 --reps 3 --mb 256 || { echo "HANG/FAIL: stop"; exit 1; }
echo "== pytest -m gpu"; timeout -s KILL 700 python -m pytest tests -m gpu -q 2>&1 | tail -15
echo "== bench"; timeout -s KILL 420 python bench.py > gpurun_out/s13_bench_n1.json 2> gpurun_out/s13_bench_n1.err; tail -c 1500 gpurun_out/s13_bench_n1.json; tail -3 gpurun_out/s13_bench_n1.err
echo "== 65536 launch list"
timeout -s KILL 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 12 --csv \
    --log-file gpurun_out/s13_solo65536_launches.csv python tools/kbench.py --n 65536 --mb 2048 --reps 1 --no-check > /dev/null 2>&1
grep -v "^==" gpurun_out/s13_solo65536_launches.csv | awk -F'","' '{print $5, $(NF-2), $(NF)}' | cut -c1-140 | tail -12
for tool in memcheck synccheck; do
  echo "== sanitizer $tool"
  timeout -s KILL 300 compute-sanitizer --tool $tool python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
      -k "(solo65536 and (matches and 149 or window and 149)) or mixes_both" > gpurun_out/s13_sanitizer_$tool.txt 2>&1
  tail -4 gpurun_out/s13_sanitizer_$tool.txt
done
} > gpurun_out/step13.txt 2>&1
cat gpurun_out/step13.txt
