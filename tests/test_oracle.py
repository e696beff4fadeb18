"""CPU tests: the restatement (oracle/rpf_oracle.c) against the analytic
known-answer tests of SURVEY.md section 4, against numpy float64, and against
the reference's own src/datastore.cxx (oracle/_ref, stand-in FFT)."""
import numpy as np
import pytest

import oracle_lib as ol

RNG = np.random.default_rng(0x5EED)


def rand_stream(nbytes, seed=1):
    return np.random.default_rng(seed).integers(0, 256, size=nbytes, dtype=np.uint8)


def relmax(a, b):
    return np.max(np.abs(a - b)) / np.max(np.abs(b))


@pytest.mark.parametrize("N", [2, 8, 64, 512, 4096])
def test_zero_signal(N):
    pwr, done = ol.oracle_run_stream(N, 3, np.full(2 * N * 3, 127, np.uint8))
    assert done == 3
    assert np.all(pwr == 0.0)


@pytest.mark.parametrize("N", [8, 512, 4096])
def test_dc_lands_in_middle_bin(N):
    a = 40
    s = np.full(2 * N, 127, np.uint8)
    s[0::2] = 127 + a
    pwr, _ = ol.oracle_run_stream(N, 1, s)
    assert np.argmax(pwr) == N // 2
    assert pwr[N // 2] == pytest.approx((a * N) ** 2, rel=1e-12)
    rest = np.delete(pwr, N // 2)
    assert np.max(rest) <= 1e-18 * pwr[N // 2]


def test_tone_index_and_nyquist():
    N, k, A = 512, 37, 100.0
    n = np.arange(N)
    s = np.empty(2 * N, np.uint8)
    s[0::2] = np.round(127 + A * np.cos(2 * np.pi * k * n / N))
    s[1::2] = np.round(127 + A * np.sin(2 * np.pi * k * n / N))
    pwr, _ = ol.oracle_run_stream(N, 1, s)
    assert np.argmax(pwr) == N // 2 + k          # index 293
    assert pwr[N // 2 + k] == pytest.approx((A * N) ** 2, rel=2e-2)
    s = np.full(2 * N, 127, np.uint8)
    s[0::2] = 127 + 50 * (1 - 2 * (n % 2))       # I alternates 127 +- 50
    pwr, _ = ol.oracle_run_stream(N, 1, s)
    assert np.argmax(pwr) == 0


@pytest.mark.parametrize("N", [16, 500, 512, 4096])
def test_parseval(N):
    R = 5
    s = rand_stream(2 * N * R, seed=N)
    pwr, done = ol.oracle_run_stream(N, R, s)
    x = s.astype(np.int64) - 127
    assert done == R
    assert pwr.sum() == pytest.approx(N * float((x * x).sum()), rel=1e-12)


@pytest.mark.parametrize("N,R", [(2, 7), (6, 3), (500, 2), (512, 100), (4096, 9), (65536, 2)])
def test_matches_numpy_float64(N, R):
    s = rand_stream(2 * N * R, seed=N + R)
    pwr, done = ol.oracle_run_stream(N, R, s)
    assert done == R
    assert relmax(pwr, ol.numpy_power(N, s, R)) < 1e-12


def test_window_matches_numpy():
    N, R = 4096, 4
    w = (0.5 - 0.5 * np.cos(2 * np.pi * np.arange(N) / N)).astype(np.float32)
    s = rand_stream(2 * N * R, seed=9)
    pwr, _ = ol.oracle_run_stream(N, R, s, window=w)
    assert relmax(pwr, ol.numpy_power(N, s, R, window=w)) < 1e-12


def test_f32_flavour_close_to_f64():
    N, R = 4096, 64
    s = rand_stream(2 * N * R, seed=3)
    p64, _ = ol.oracle_run_stream(N, R, s, fft_kind=0)
    p32, _ = ol.oracle_run_stream(N, R, s, fft_kind=1)
    assert np.max(np.abs(p32 - p64) / p64) < 1e-5     # measured ~1e-7


def test_truncation_and_surplus():
    """Config 1: N=512, repeats=100 -> one 114,688-byte read (112 spectra) of which
    100 are used (src/acquisition.cxx:295, src/datastore.cxx:67)."""
    N, R = 512, 100
    assert ol.oracle().rpf_oracle_plan_buf_length(N, R, 1638400, 0) == 114688
    sizes = ol.oracle_read_sizes(N, R, 114688)
    assert list(sizes) == [114688]
    s = rand_stream(114688, seed=5)
    pwr, done = ol.oracle_run_stream(N, R, s, buf_length=114688)
    assert done == 100
    s2 = s.copy()
    s2[2 * N * R:] = 0                                  # surplus bytes have no effect
    pwr2, _ = ol.oracle_run_stream(N, R, s2, buf_length=114688)
    assert np.array_equal(pwr, pwr2)
    assert relmax(pwr, ol.numpy_power(N, s, R)) < 1e-12


def test_straddle_equals_contiguous():
    """N=65536 with 1,638,400-byte buffers: 12.5 spectra per buffer (fact 4)."""
    N, R = 65536, 3
    s = rand_stream(2 * N * R, seed=11)
    w = ol.OracleWorker(N, R)
    for off in range(0, s.size, 16384 * 5 + 0):       # deliberately not spectrum aligned
        w.push(s[off:off + 16384 * 5])
    assert w.repeats_done == R
    one, _ = ol.oracle_run_stream(N, R, s, buf_length=2 * N * R)
    assert np.array_equal(w.pwr, one)


def test_partial_spectrum_dropped_and_carry():
    N = 64
    s = rand_stream(2 * N * 2 + 2 * 10, seed=2)         # 2 spectra + 10 samples
    w = ol.OracleWorker(N, 100)
    w.push(s)
    assert w.repeats_done == 2 and w.fft_pointer == 10
    w.reset(100)
    assert w.repeats_done == 0 and w.fft_pointer == 0 and np.all(w.pwr == 0)


def test_read_sizes_rule():
    # 4096 bins, 1e6 repeats, default buffer: 5000 full reads
    sizes = ol.oracle_read_sizes(4096, 1000000, 1638400)
    assert len(sizes) == 5000 and np.all(sizes == 1638400)
    # remainder rounded up to a multiple of 16384 (src/acquisition.cxx:295)
    sizes = ol.oracle_read_sizes(512, 1700, 1638400)   # 1,740,800 bytes
    assert list(sizes) == [1638400, 114688]


def test_write_data_numeric():
    import ctypes as C
    N = 8
    pwr = np.arange(1, N + 1, dtype=np.float64) * 1e6
    f = np.zeros(N)
    v = np.zeros(N)
    p = pwr.copy()
    ol.oracle().rpf_oracle_write_data(N, p.ctypes.data_as(C.POINTER(C.c_double)), 10, 100e6, 2000000, 0, None,
                                      f.ctypes.data_as(C.POINTER(C.c_double)), v.ctypes.data_as(C.POINTER(C.c_double)))
    assert p[N // 2] == (pwr[N // 2 - 1] + pwr[N // 2 + 1]) / 2
    assert f[0] == 100e6 - 1e6 and f[N // 2] == 100e6
    assert v[0] == pytest.approx(10 * np.log10(pwr[0] / 10 / N / 2e6))


# ---------------------------------------------------------------------------
# The restatement against the reference's own worker (oracle/_ref).
# ---------------------------------------------------------------------------
needs_ref = pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref not built (reference sources absent)")


@needs_ref
@pytest.mark.parametrize("N,R,buf", [(512, 100, None), (4096, 450, 1638400), (65536, 30, 1638400),
                                     (500, 40, None), (2, 5, 16384), (4096, 7, 16384)])
def test_restatement_matches_reference_worker(N, R, buf):
    s = rand_stream(min(2 * N * R + 16384, 1 << 22), seed=N ^ R)
    wrap = s.size < 2 * N * R + 16384
    pr, done_r, hist, reads = ol.ref_run(N, R, s, buf_length=buf, wrap=wrap)
    po, done_o = ol.oracle_run_stream(N, R, s, buf_length=buf, wrap=wrap, fft_kind=0)
    assert done_r == done_o == R
    # only the DFT routine differs (float32 stand-in vs float64): ~1e-7
    assert relmax(pr, po) < 2e-6
    assert np.argmax(pr) == np.argmax(po)
    assert sum(hist) == reads


@needs_ref
def test_reference_worker_window_and_kat():
    N, R = 4096, 20
    w = (0.5 - 0.5 * np.cos(2 * np.pi * np.arange(N) / N)).astype(np.float32)
    s = rand_stream(2 * N * R, seed=77)
    pr, done, _, _ = ol.ref_run(N, R, s, window=w, buf_length=16384 * 10)
    po, _ = ol.oracle_run_stream(N, R, s, window=w, buf_length=16384 * 10)
    assert done == R and relmax(pr, po) < 2e-6
    # tone KAT through the reference's own loop: bin N/2 + k
    k, A = 300, 90.0
    n = np.arange(N)
    t = np.empty(2 * N, np.uint8)
    t[0::2] = np.round(127 + A * np.cos(2 * np.pi * k * n / N))
    t[1::2] = np.round(127 + A * np.sin(2 * np.pi * k * n / N))
    pr, _, _, _ = ol.ref_run(N, 1, t, buf_length=16384)
    assert np.argmax(pr) == N // 2 + k
