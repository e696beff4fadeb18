#!/usr/bin/env python
"""Generates tests/golden/ref_text.json: the text (stdout) and summary (stderr) that the REFERENCE's own
acquisition layer prints -- src/acquisition.cxx (AuxData, Plan, Acquisition::run, print_summary, write_data),
src/device.cxx, src/datastore.cxx, src/interrupts.cxx compiled unmodified from /root/reference into
oracle/_ref/ref_rtl_power behind a fake dongle and the stand-in FFT (oracle/Makefile) -- for a few small cases,
so that the C++ host side of this repository can be compared with it on the GPU box, where /root/reference does
not exist.  Run in the build container:  python tests/golden/make_ref_text_golden.py
The byte streams are not stored: they come from the repository's counter-based generator (seed, mode below).
"""
import importlib
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.path.join(ROOT, "oracle", "_ref", "ref_rtl_power")

# name: (N, repeats, seed, mode, stream_bytes, reference options, rpf_power_synth options)
CASES = {
    "cfg1_512x100": (512, 100, 0x5EED + 512, 1, 114688, [], []),
    "n1024_linear_baseline_window": (1024, 40, 0x5EED + 1024, 1, 16384 * 5,
                                     ["--linear", "-w", "@hann", "-B", "-"], ["-w", "@hann", "-B", "-"]),
    "n64_three_hops": (64, 500, 77, 2, 16384 * 4, ["--hops", "3"], ["--hops", "3"]),
    "n2048_straddle": (2048, 30, 31, 0, 16384 * 9, ["--buf-length", str(16384 * 3)], []),   # 3 reads of 49,152 bytes
    "n500_bluestein": (500, 40, 0x5EED + 500, 1, 49152, [], []),
}


def aux_text(kind, n):
    if kind == "@hann":
        return "".join("%d %.9g\n" % (i, 0.5 - 0.5 * np.cos(2 * np.pi * i / n)) for i in range(n))
    return "# baseline\n" + "".join("%.17g\n" % (-1.0 + 2.0 * i / n) for i in range(n))


def strip_times(text):
    return "\n".join(ln for ln in text.splitlines() if not ln.startswith("# Acquisition ")) + "\n"


def main():
    pkg = importlib.import_module("rtl-power-fftw_b200")
    out = {}
    for name, (N, rep, seed, mode, nbytes, ropts, _) in CASES.items():
        with tempfile.TemporaryDirectory() as td:
            stream = pkg.synth_fill_host(nbytes, 0, seed=seed, mode=mode)
            sp = os.path.join(td, "s.bin")
            stream.tofile(sp)
            args = []
            for o in ropts:
                if o.startswith("@"):
                    fp = os.path.join(td, o[1:] + ".txt")
                    open(fp, "w").write(aux_text(o, N))
                    args.append(fp)
                else:
                    args.append(o)
            stdin = aux_text("@ramp", N) if "-" in args else None           # "-B -": baseline from stdin
            r = subprocess.run([REF, sp, str(N), str(rep)] + args, input=stdin, capture_output=True, text=True, check=True)
            keep = [ln for ln in r.stderr.splitlines()
                    if ln.startswith(("Actual number", "Number of successful", "Effective integration", "Buffer queue"))]
            out[name] = {"stdout": strip_times(r.stdout), "stderr": keep}
            print(name, len(r.stdout.splitlines()), "lines;", keep[-1])
    json.dump(out, open(os.path.join(HERE, "ref_text.json"), "w"), indent=0, sort_keys=True)


if __name__ == "__main__":
    main()
