#!/usr/bin/env python
"""Generates tests/golden/*.npz -- small input/output vectors for the hot path.

Run in the build container (where /root/reference exists):  python tests/golden/make_golden.py

Each fixture holds the exact input bytes and
  pwr_ref    : output of the reference's own worker, src/datastore.cxx:48-96 compiled unmodified
               into oracle/_ref (FFT = stand-in float32 Stockham, libfftw3f being absent), driven
               through the reference's buffer queues with the reference's read sizes;
  pwr_f64    : the float64 restatement (oracle/rpf_oracle.c) on the same bytes;
so the GPU path can be checked on the GPU box, where /root/reference does not exist.
Inputs come from numpy's PCG64 (not the repo's generator), seeds below.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as ol  # noqa: E402

CASES = [
    # name,           N,     repeats, window, buf_length, kind
    ("cfg1_512x100",  512,   100,     False,  None,       "uniform"),     # BASELINE config 1
    ("n4096x24",      4096,  24,      False,  65536,      "gauss"),
    ("n4096x12_hann", 4096,  12,      True,   49152,      "gauss"),       # config 3 shape
    ("n65536x3",      65536, 3,       False,  1638400 // 4, "gauss"),     # config 4 shape: straddling buffers
    ("n64x50_tone",   64,    50,      False,  16384,      "tone"),
    ("n2048x9",       2048,  9,       False,  16384,      "uniform"),
    # one read of 120 spectra: the launch goes to the one-CTA-per-spectrum kernel (>= 96 spectra; hand-off through
    # tensor memory at 8192 bins) -- appended last so that the seeds of the fixtures above do not move
    ("n8192x120_one_read", 8192, 120, False,  2 * 8192 * 120, "gauss"),
]


def make_stream(kind, nbytes, N, seed):
    rng = np.random.default_rng(seed)
    if kind == "uniform":
        return rng.integers(0, 256, nbytes, dtype=np.uint8)
    if kind == "gauss":
        return np.clip(np.round(rng.normal(127.4, 30.0, nbytes)), 0, 255).astype(np.uint8)
    n = np.arange(nbytes // 2)
    k = N // 4 + 3
    s = np.empty(nbytes, np.float64)
    s[0::2] = 127 + 90 * np.cos(2 * np.pi * k * n / N) + rng.normal(0, 4, n.size)
    s[1::2] = 127 + 90 * np.sin(2 * np.pi * k * n / N) + rng.normal(0, 4, n.size)
    return np.clip(np.round(s), 0, 255).astype(np.uint8)


def main():
    assert ol.have_ref(), "oracle/_ref not built: needs /root/reference"
    for i, (name, N, R, win, buf, kind) in enumerate(CASES):
        if buf is None:
            buf = int(ol.oracle().rpf_oracle_plan_buf_length(N, R, 1638400, 0))
        sizes = ol.oracle_read_sizes(N, R, buf)
        stream = make_stream(kind, int(sizes.sum()), N, 1000 + i)
        w = (0.5 - 0.5 * np.cos(2 * np.pi * np.arange(N) / N)).astype(np.float32) if win else None
        pr, done_r, hist, reads = ol.ref_run(N, R, stream, window=w, buf_length=buf)
        po, done_o = ol.oracle_run_stream(N, R, stream, window=w, buf_length=buf)
        assert done_r == done_o == R and reads == len(sizes)
        rel = float(np.max(np.abs(pr - po)) / np.max(po))
        print("%-16s N=%6d R=%4d reads=%3d  ref(f32 stand-in) vs f64 restatement: %.2e" % (name, N, R, reads, rel))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), stream=stream, N=N, repeats=R, buf_length=buf,
                            window=(w if win else np.zeros(0, np.float32)), pwr_ref=pr, pwr_f64=po,
                            repeats_done=done_r, read_sizes=sizes)


if __name__ == "__main__":
    main()
