#!/usr/bin/env python
"""Measures every single-GPU BASELINE.json configuration with a device-resident stream (bounded
sample of the long ones), CUDA-event timed, and writes one JSON document.  Not the bench line --
bench.py reports config 2; this is the table in DESIGN.md."""
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
CONFIGS = [
    # name, N, spectra in the config, spectra measured per launch, window
    ("cfg1: 512 bins, 100 spectra (one 114,688-byte read)", 512, 100, 100, False),
    ("cfg2: 4096 bins, 1e6 spectra", 4096, 1000000, 1000000, False),
    ("cfg3: 4096 bins, Hann window, 1e7 spectra (measured on 1e6)", 4096, 10000000, 1000000, True),
    ("cfg4: 65536 bins, 1e5 spectra (measured on 32768)", 65536, 100000, 32768, False),
]


def main():
    import torch
    pkg = importlib.import_module("rtl-power-fftw_b200")
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    out = []
    for name, N, total, nspec, win in CONFIGS:
        nbytes = 2 * N * nspec
        dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        s = torch.cuda.Stream()
        w = (0.5 - 0.5 * np.cos(2 * np.pi * np.arange(N) / N)).astype(np.float32) if win else None
        with torch.cuda.stream(s):
            pkg.synth_fill_device(dev.data_ptr(), nbytes, 0, seed=0x5EED + 2, mode=1, stream=s.cuda_stream)
            ctx = pkg.Context(N, window=w)
            ctx.set_stream(s.cuda_stream)
            pwr = torch.zeros(N, dtype=torch.float64, device="cuda")
            ts = []
            for i in range(8):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ctx.reset(nspec)
                ctx.push_device(dev.data_ptr(), nbytes)
                ctx.finish_device(pwr.data_ptr())
                e1.record()
                e1.synchronize()
                if i >= 3:
                    ts.append(e0.elapsed_time(e1))
            ctx.close()
        ms = float(np.median(ts))
        rate = nspec * N / (ms * 1e-3)
        out.append({"config": name, "bins": N, "window": win, "spectra_per_launch": nspec, "ms_per_integration": ms,
                    "Msamples_per_s": rate / 1e6, "algorithmic_GBps": 2 * rate / 1e9, "frac_of_hbm_roofline": 2 * rate / 1e9 / peak,
                    "plan": None})
        del dev
        print(out[-1], flush=True)
    json.dump({"hbm_peak_GBps": peak, "configs": out}, open(os.path.join(ROOT, "gpurun_out", "configs.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
