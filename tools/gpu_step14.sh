#!/bin/bash
# short pushes read in place from page-locked memory: tests + config 1 latency
mkdir -p gpurun_out
{
echo "== tests"; timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_host.py tests/test_golden.py -m gpu -q -k "page_locked or config1 or host or golden or straddl or repeats" 2>&1 | tail -5
echo "== config 1"; timeout -s KILL 120 python - <<'PY'
import sys, importlib, json
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import bench
pkg = importlib.import_module('rtl-power-fftw_b200')
for _ in range(3):
    r = bench.bench_config1(pkg, 0, reps=1000)
    print(json.dumps({k: r[k] for k in ("us_per_integration", "Msamples_per_s", "parity", "kernel")}))
PY
} > gpurun_out/step14.txt 2>&1
cat gpurun_out/step14.txt
