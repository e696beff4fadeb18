#!/bin/bash
# bench.py prints exactly one line on stdout; one --set full capture of the tensor-memory hand-off kernel (8192 bins)
mkdir -p gpurun_out
{
echo "== bench stdout"; timeout -s KILL 200 python bench.py --steps 5 --no-configs > gpurun_out/s20_bench.json 2> gpurun_out/s20_bench.err
wc -l gpurun_out/s20_bench.json; head -c 300 gpurun_out/s20_bench.json; echo; tail -2 gpurun_out/s20_bench.err
echo "== ncu full 8192"
timeout -s KILL 200 ncu --set full --import-source on --clock-control none -k regex:solo65536 -c 1 -o gpurun_out/solo8192_full -f python tools/kbench.py --n 8192 --mb 1024 --reps 1 --no-check > /dev/null 2>&1
ls -la gpurun_out/solo8192_full.ncu-rep
} > gpurun_out/step20.txt 2>&1
cat gpurun_out/step20.txt
