#!/bin/bash
# determinism (10 bit-identical runs of a 192 MiB launch) for the solo kernel at every outer radix
mkdir -p gpurun_out
timeout -s KILL 120 python -m pytest tests/test_gpu_host.py -m gpu -q -k "nondeterminism and (8192 or 16384 or 32768 or 65536)" > gpurun_out/step22.txt 2>&1
tail -4 gpurun_out/step22.txt
