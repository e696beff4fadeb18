"""Committed golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py from the
reference's own worker loop, oracle/_ref).  CPU: the oracle reproduces them.  GPU: the CUDA path,
through the C ABI, fed the same bytes in the same read sizes, matches them."""
import glob
import os

import numpy as np
import pytest

import oracle_lib as ol

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = sorted(glob.glob(os.path.join(HERE, "golden", "*.npz")))


def load(path):
    z = np.load(path)
    w = z["window"]
    return dict(N=int(z["N"]), R=int(z["repeats"]), buf=int(z["buf_length"]), stream=z["stream"],
                window=(w if w.size else None), pwr_ref=z["pwr_ref"], pwr_f64=z["pwr_f64"],
                done=int(z["repeats_done"]), sizes=z["read_sizes"])


def test_fixtures_present():
    assert len(FIXTURES) >= 6


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[:-4] for p in FIXTURES])
def test_oracle_reproduces_golden(path):
    f = load(path)
    po, done = ol.oracle_run_stream(f["N"], f["R"], f["stream"], window=f["window"], buf_length=f["buf"])
    assert done == f["done"]
    assert np.array_equal(po, f["pwr_f64"])                                   # the restatement is deterministic
    assert np.max(np.abs(po - f["pwr_ref"])) / np.max(f["pwr_ref"]) < 2e-6    # vs the reference's loop (f32 FFT)
    assert np.argmax(po) == np.argmax(f["pwr_ref"])


@pytest.mark.gpu
@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[:-4] for p in FIXTURES])
def test_gpu_matches_golden(path):
    from rpf import Context
    f = load(path)
    with Context(f["N"], window=f["window"], buf_length=f["buf"]) as c:
        c.reset(f["R"])
        off = 0
        for sz in f["sizes"]:                        # the reference producer's read sizes
            c.push(f["stream"][off:off + int(sz)])
            off += int(sz)
        pwr, done = c.finish()
    assert done == f["done"]
    scale = np.max(f["pwr_ref"])
    assert np.max(np.abs(pwr - f["pwr_ref"])) / scale < 1e-5                  # vs the reference's own worker
    assert np.max(np.abs(pwr - f["pwr_f64"])) / scale < 1e-5                  # vs float64 truth
    assert np.argmax(pwr) == np.argmax(f["pwr_ref"])
