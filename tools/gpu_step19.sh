#!/bin/bash
# 2-GPU validation: multi-device / multi-rank tests and the 2-GPU bench line
mkdir -p gpurun_out
{
echo "== sanity of the last host-side change (no scratch / no persisting limit when the hand-off is in tensor memory)"
timeout -s KILL 60 python tools/kbench.py --n 8192 --reps 3 --mb 1024 || { echo "HANG/FAIL: stop"; exit 1; }
timeout -s KILL 60 python tools/kbench.py --n 32768 --reps 3 --mb 1024
timeout -s KILL 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "solo or mixes" 2>&1 | tail -2
echo "== multi-GPU tests"; timeout -s KILL 400 python -m pytest tests/test_gpu_multidevice.py tests/test_gpu_multirank.py -m gpu -q 2>&1 | tail -6
echo "== bench --gpus 2"
timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/s19_bench_n2.json 2> gpurun_out/s19_bench_n2.err
tail -c 1200 gpurun_out/s19_bench_n2.json; tail -3 gpurun_out/s19_bench_n2.err
} > gpurun_out/step19.txt 2>&1
cat gpurun_out/step19.txt
