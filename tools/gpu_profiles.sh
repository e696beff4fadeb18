#!/bin/bash
# Regenerates the evidence kept under profiles/: bench lines, ncu launch list of the bench command, one
# --set full capture of the headline kernel (details + selected raw counters + DRAM traffic of a whole
# bench-sized launch), the config sweep.  Run under gpurun; copy gpurun_out/prof_* into profiles/ afterwards.
set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/prof_bench_n1.json 2> gpurun_out/prof_bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/prof_bench_reference_n1.json 2>> gpurun_out/prof_bench_n1.err
python tools/config_sweep.py > gpurun_out/prof_configs.log 2>&1
# launch list of the bench command (times are cold-cache and serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/prof_launches_bench_steps2.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/prof_launches.log 2>&1
# full capture of the headline kernel on a 1 GiB stream, and DRAM traffic of a bench-sized (8.19 GB) launch
ncu --set full --clock-control none --import-source on -k regex:pipe4096 -s 1 -c 1 -f -o gpurun_out/prof_pipe4096 \
    python tools/kbench.py --n 4096 --mb 1024 --reps 1 --no-check > gpurun_out/prof_ncu_full.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:pipe4096 -s 1 -c 1 --csv \
    --log-file gpurun_out/prof_traffic_bench_launch.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/prof_traffic.log 2>&1
ls -la gpurun_out
tail -c 600 gpurun_out/prof_bench_n1.json
