#!/bin/bash
# Regenerates the evidence kept under profiles/ (round 2): bench lines, ncu launch list of the bench command, one
# --set full capture of the headline kernel and of the tensor-core variant, DRAM traffic of a whole bench-sized
# launch, the 65536-bin launch list with DRAM bytes, sanitizer runs.  Run under gpurun; summaries are then made
# from gpurun_out/prof_* with tools/ncu_stalls.py and copied into profiles/.
set -x
mkdir -p gpurun_out
timeout -s KILL 400 python bench.py > gpurun_out/prof_bench_n1.json 2> gpurun_out/prof_bench_n1.err
timeout -s KILL 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/prof_bench_reference_n1.json 2>> gpurun_out/prof_bench_n1.err
# launch list of the bench command (times are cold-cache and serialised: compare shares)
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/prof_launches_bench_steps2.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-configs --no-parity > gpurun_out/prof_launches.log 2>&1
# full capture of the headline kernel on a 1 GiB stream, and DRAM traffic of a bench-sized (8.19 GB) launch
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:pipe4096_kernel -s 1 -c 1 -f -o gpurun_out/prof_pipe4096 \
    python tools/kbench.py --n 4096 --mb 1024 --reps 1 --no-check > gpurun_out/prof_ncu_full.log 2>&1
timeout -s KILL 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:pipe4096_kernel -s 3 -c 1 --csv \
    --log-file gpurun_out/prof_traffic_bench_launch.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-configs --no-parity > gpurun_out/prof_traffic.log 2>&1
# 65536 bins: every launch of one push of 2048 spectra (4 batches) with duration and DRAM bytes
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 40 --csv \
    --log-file gpurun_out/prof_fourstep_launches.csv python tools/kbench.py --n 65536 --mb 256 --reps 1 --no-check > gpurun_out/prof_fourstep.log 2>&1
# config 1 (512 x 100 through rpf_push): the launches of a few integrations
timeout -s KILL 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file gpurun_out/prof_config1_launches.csv \
    python -c "
import sys, importlib
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
pkg = importlib.import_module('rtl-power-fftw_b200')
s = pkg.synth_fill_host(114688, 0, seed=0x5EED + 1, mode=1)
with pkg.Context(512) as c:
    for _ in range(6):
        c.reset(100); c.push(s); c.finish()
" > gpurun_out/prof_config1.log 2>&1
# the small-bin pipeline: one full capture each of pipen_kernel<2048> (forced) and pipen_kernel<512>
timeout -s KILL 200 ncu --set full --clock-control none --import-source on -k regex:pipen_kernel -s 1 -c 1 -f -o gpurun_out/prof_pipen2048 \
    python tools/kbench.py --n 2048 --mb 1024 --reps 1 --no-check --flags 8 > gpurun_out/prof_pipen2048.log 2>&1
timeout -s KILL 200 ncu --set full --clock-control none --import-source on -k regex:pipen_kernel -s 1 -c 1 -f -o gpurun_out/prof_pipen512 \
    python tools/kbench.py --n 512 --mb 1024 --reps 1 --no-check > gpurun_out/prof_pipen512.log 2>&1
# sanitizers over the round-2 kernels (pipen, TMA outer pass, cluster kernel, tensor-core variant, Bluestein, multi-batch four-step)
for tool in memcheck synccheck racecheck; do
  timeout -s KILL 500 compute-sanitizer --tool $tool python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
      -k "tc_pass1 or every_even_bin_count or config1 or tail_shapes or (tma_outer and 8192) or (cluster65536 and 17)" > gpurun_out/sanitizer_$tool.txt 2>&1
  tail -5 gpurun_out/sanitizer_$tool.txt
done
ls -la gpurun_out
tail -c 600 gpurun_out/prof_bench_n1.json
