#!/bin/bash
# final 1-GPU validation of the round: whole GPU suite, smoke, bench line (+ reference arm), launch lists, sanitizers
mkdir -p gpurun_out
{
echo "== smoke"; timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== pytest -m gpu"; timeout -s KILL 900 python -m pytest tests -m gpu -q 2>&1 | tail -12
echo "== bench"; timeout -s KILL 420 python bench.py > gpurun_out/s18_bench_n1.json 2> gpurun_out/s18_bench_n1.err; tail -c 600 gpurun_out/s18_bench_n1.json; tail -3 gpurun_out/s18_bench_n1.err
echo "== bench reference arm"; timeout -s KILL 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/s18_bench_reference_n1.json 2>> gpurun_out/s18_bench_n1.err; tail -c 400 gpurun_out/s18_bench_reference_n1.json
echo "== launch list of the bench command"
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s18_launches_bench_steps2.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-configs --no-parity > /dev/null 2>&1
grep -v "^==" gpurun_out/s18_launches_bench_steps2.csv | awk -F'","' '{print substr($5,1,60), $(NF)}' | sort | uniq -c | sort -rn | head -8
echo "== 8192 launch list (solo, tensor-memory hand-off)"
timeout -s KILL 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 6 --csv \
    --log-file gpurun_out/s18_solo8192_launches.csv python tools/kbench.py --n 8192 --mb 2048 --reps 1 --no-check > /dev/null 2>&1
grep -v "^==" gpurun_out/s18_solo8192_launches.csv | awk -F'","' '{print substr($5,1,50), $(NF-2), $(NF)}' | tail -9
echo "== config 1 launch list"
timeout -s KILL 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file gpurun_out/s18_config1_launches.csv \
    python -c "
import sys, importlib
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
pkg = importlib.import_module('rtl-power-fftw_b200')
s = pkg.synth_fill_host(114688, 0, seed=0x5EED + 1, mode=1)
with pkg.Context(512) as c:
    for _ in range(6):
        c.reset(100); c.push(s); c.finish()
" > /dev/null 2>&1
grep -v "^==" gpurun_out/s18_config1_launches.csv | awk -F'","' '{print substr($5,1,60), $(NF)}' | tail -6
for tool in memcheck synccheck; do
  echo "== sanitizer $tool"
  timeout -s KILL 400 compute-sanitizer --tool $tool python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
      -k "(other_outer_radices and (149 and not True)) or (solo65536 and matches and 149) or mixes_both or page_locked" > gpurun_out/s18_sanitizer_$tool.txt 2>&1
  tail -4 gpurun_out/s18_sanitizer_$tool.txt
done
} > gpurun_out/step18.txt 2>&1
cat gpurun_out/step18.txt
