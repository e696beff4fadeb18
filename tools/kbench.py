#!/usr/bin/env python
"""Kernel micro-benchmark: device-resident stream, CUDA-event timing of rpf_push_device
(one fused launch), plus a quick parity check against the oracle.  Development tool."""
import argparse
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, nargs="+", default=[4096])
    ap.add_argument("--mb", type=int, default=2048, help="stream size in MiB")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--window", action="store_true")
    ap.add_argument("--no-check", action="store_true")
    ap.add_argument("--lib", default=None, help="alternative librpf_b200.so (A/B builds under variants/)")
    ap.add_argument("--scratch-mib", type=int, default=0)
    ap.add_argument("--flags", type=int, default=0, help="rpf_config.flags (1 = RPF_FLAG_TC_PASS1)")
    a = ap.parse_args()
    import torch
    pkg = importlib.import_module("rtl-power-fftw_b200")
    if a.lib:
        pkg.capi.LIB_PATH = os.path.abspath(a.lib)       # before the first lib() call
    import oracle_lib as ol
    peak = 6553.9
    for N in a.n:
        nspec = (a.mb << 20) // (2 * N)
        nbytes = 2 * N * nspec
        dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        s = torch.cuda.Stream()
        w = (0.5 - 0.5 * np.cos(2 * np.pi * np.arange(N) / N)).astype(np.float32) if a.window else None
        with torch.cuda.stream(s):
            pkg.synth_fill_device(dev.data_ptr(), nbytes, 0, seed=1, mode=1, stream=s.cuda_stream)
            ctx = pkg.Context(N, window=w, flags=a.flags, scratch_mib=a.scratch_mib)
            ctx.set_stream(s.cuda_stream)
            ts = []
            for i in range(a.reps + 2):
                ctx.reset(nspec)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ctx.push_device(dev.data_ptr(), nbytes)
                e1.record()
                e1.synchronize()
                if i >= 2:
                    ts.append(e0.elapsed_time(e1))
            pwr, done = ctx.finish()
        ms = float(np.median(ts))
        rate = nspec * N / (ms * 1e-3)
        msg = "N=%6d spectra=%8d  %.3f ms  %.1f Gsamples/s  %.1f GB/s  %.1f%% of %.0f GB/s" % (
            N, nspec, ms, rate / 1e9, 2 * rate / 1e9, 100 * 2 * rate / 1e9 / peak, peak)
        if not a.no_check:
            R = min(nspec, 24 if N <= 4096 else 3)
            host = pkg.synth_fill_host(2 * N * R, 0, seed=1, mode=1)
            ctx.reset(R)
            ctx.push(host)
            p, _ = ctx.finish()
            po, _ = ol.oracle_run_stream(N, R, host, window=w, buf_length=host.size)
            msg += "  parity %.2e" % float(np.max(np.abs(p - po) / po))
        ctx.close()
        print(msg, flush=True)
        del dev


if __name__ == "__main__":
    main()
