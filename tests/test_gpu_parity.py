"""GPU parity tests: the CUDA path, called through the C ABI (include/rpf_b200.h),
against the CPU oracle on identical bytes.

Tolerance (BASELINE.json north_star): 1e-5 relative, float32 path vs the
float64 oracle; bin indices exact.  For noise-like input the bound is applied
per bin; for tonal input relative to the largest bin."""
import numpy as np
import pytest

import oracle_lib as ol
from rpf import Context, RpfError, pkg

pytestmark = pytest.mark.gpu
TOL = 1e-5


def stream(nbytes, seed=1, mode=0):
    return pkg.synth_fill_host(nbytes, 0, seed=seed, mode=mode)


def run_gpu(N, repeats, chunks, window=None, flags=0):
    with Context(N, window=window, flags=flags) as c:
        c.reset(repeats)
        for ch in chunks:
            c.push(ch)
        return c.finish()


def perbin(a, b):
    return np.max(np.abs(a - b) / b)


def relmax(a, b):
    return np.max(np.abs(a - b)) / np.max(np.abs(b))


@pytest.mark.parametrize("N", [2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 65536])
def test_random_bytes_all_sizes(N):
    R = max(3, min(40, (1 << 19) // N))
    s = stream(2 * N * R, seed=N)
    pg, done = run_gpu(N, R, [s])
    po, done_o = ol.oracle_run_stream(N, R, s, buf_length=s.size)
    assert done == done_o == R
    assert perbin(pg, po) < TOL
    assert np.argmax(pg) == np.argmax(po)


@pytest.mark.parametrize("N,R", [(6, 40), (10, 33), (100, 30), (130, 30), (500, 25), (1000, 12), (3000, 4), (8190, 2),
                                 (10000, 5), (12346, 3), (100000, 2), (131072, 3), (262144, 2)])
def test_every_even_bin_count(N, R):
    """The reference only forces N even (src/params.cxx:152-155; man page: -b 500) and FFTW takes any length:
    O(N^2) kernel up to 128 bins, Bluestein above, plain FFT passes for powers of two beyond 65536."""
    s = stream(2 * N * R + 2 * (N // 2), seed=N)      # + half a spectrum
    w = (0.5 - 0.5 * np.cos(2 * np.pi * np.arange(N) / N)).astype(np.float32)
    want = "gfft" if N & (N - 1) == 0 else ("direct" if N <= 128 else "bluestein")
    for window in (None, w):
        with Context(N, window=window) as c:
            assert c.plan_name == want
            c.reset(10 ** 6)
            cut = (s.size // 6) * 2
            c.push(s[:cut])
            c.push(s[cut:])
            pg, done = c.finish()
            assert done == R and c.fft_pointer == N // 2
        if N <= 3000:
            po, _ = ol.oracle_run_stream(N, R, s, window=window, buf_length=s.size)
        else:                                             # the oracle's own O(N^2) DFT is too slow here: numpy float64
            po = ol.numpy_power(N, s, R, window=window)
        assert relmax(pg, po) < TOL
        assert np.max(np.abs(pg - po) / po) < 1e-4 and np.argmax(pg) == np.argmax(po)


def test_bluestein_batches_kats_and_linearity():
    """More than one pass over the Bluestein ping-pong buffers in a single push (batch = 128 MiB / 8 M bytes), the
    analytic known answers at a non-power-of-two size, and float64 linearity over repeated pushes."""
    N = 500                                               # M = 1024 -> 16384 spectra per batch
    z = np.full(2 * N * 3, 127, np.uint8)
    with Context(N) as c:
        c.reset(3)
        c.push(z)
        p, done = c.finish()
        assert done == 3 and np.max(np.abs(p)) < 1e-3     # zero input (Bluestein is not bit-exact zero: |c| rounding)
        a, k, n = 40, 37, np.arange(N)
        d = np.empty(2 * N, np.uint8)
        d[0::2] = np.round(127 + 100 * np.cos(2 * np.pi * k * n / N))
        d[1::2] = np.round(127 + 100 * np.sin(2 * np.pi * k * n / N))
        c.reset(1)
        c.push(d)
        p, _ = c.finish()
        assert np.argmax(p) == N // 2 + k                 # tone index, fftshifted
        R = 40000                                         # 3 batches
        s = stream(2 * N * R, seed=3, mode=1)
        c.reset(R)
        l0 = c.kernel_launches
        c.push(s)
        pr, dr = c.finish()
        assert dr == R and c.kernel_launches - l0 > 2 * 13
        x = s.astype(np.int64) - 127
        assert pr.sum() == pytest.approx(N * float((x * x).sum()), rel=2e-6)          # Parseval
        po = ol.numpy_power(N, s, 2000)
        c.reset(2000)
        c.push(s)
        p2k, _ = c.finish()
        assert perbin(p2k, po) < TOL


@pytest.mark.parametrize("N", [8, 512, 4096, 65536])
def test_kat_zero_dc_nyquist_tone(N):
    z = np.full(2 * N * 2, 127, np.uint8)
    p, done = run_gpu(N, 2, [z])
    assert done == 2 and np.all(p == 0.0)                      # exact
    a = 40
    d = z[: 2 * N].copy()
    d[0::2] = 127 + a
    p, _ = run_gpu(N, 1, [d])
    assert np.argmax(p) == N // 2
    assert p[N // 2] == pytest.approx((a * N) ** 2, rel=1e-6)
    assert np.max(np.delete(p, N // 2)) <= 1e-9 * p[N // 2]
    n = np.arange(N)
    q = z[: 2 * N].copy()
    q[0::2] = 127 + 50 * (1 - 2 * (n % 2))
    p, _ = run_gpu(N, 1, [q])
    assert np.argmax(p) == 0
    k, A = N // 8 + 1, 100.0
    t = np.empty(2 * N, np.uint8)
    t[0::2] = np.round(127 + A * np.cos(2 * np.pi * k * n / N))
    t[1::2] = np.round(127 + A * np.sin(2 * np.pi * k * n / N))
    p, _ = run_gpu(N, 1, [t])
    po, _ = ol.oracle_run_stream(N, 1, t, buf_length=t.size)
    assert np.argmax(p) == N // 2 + k
    assert relmax(p, po) < TOL


@pytest.mark.parametrize("N", [16, 512, 4096, 16384])
def test_parseval(N):
    R = 6
    s = stream(2 * N * R, seed=7 + N)
    p, _ = run_gpu(N, R, [s])
    x = s.astype(np.int64) - 127
    assert p.sum() == pytest.approx(N * float((x * x).sum()), rel=1e-6)


@pytest.mark.parametrize("N,mode", [(512, 1), (4096, 1), (4096, 2), (65536, 2)])
def test_gaussian_and_tone_modes(N, mode):
    R = 12 if N < 65536 else 3
    s = stream(2 * N * R, seed=3, mode=mode)
    pg, _ = run_gpu(N, R, [s])
    po, _ = ol.oracle_run_stream(N, R, s, buf_length=s.size)
    assert relmax(pg, po) < TOL
    if mode == 1:
        assert perbin(pg, po) < TOL
    else:
        assert np.argmax(pg) == N // 2 + N // 8 == np.argmax(po)


@pytest.mark.parametrize("N", [512, 4096, 65536])
def test_window_hann(N):
    R = 5
    w = (0.5 - 0.5 * np.cos(2 * np.pi * np.arange(N) / N)).astype(np.float32)
    s = stream(2 * N * R, seed=11)
    pg, _ = run_gpu(N, R, [s], window=w)
    po, _ = ol.oracle_run_stream(N, R, s, window=w, buf_length=s.size)
    assert relmax(pg, po) < TOL
    assert np.max(np.abs(pg - po) / po) < 1e-4      # per bin, incl. the deep skirts of the DC line


def test_config1_reference_shape():
    """BASELINE config 1: 512 bins, 100 spectra, one 114,688-byte read, 12 surplus spectra."""
    N, R = 512, 100
    s = stream(114688, seed=0x5EED + 1)
    pg, done = run_gpu(N, R, [s])
    po, _ = ol.oracle_run_stream(N, R, s, buf_length=114688)
    assert done == 100 and perbin(pg, po) < TOL
    s2 = s.copy()
    s2[2 * N * R:] = 0
    pg2, _ = run_gpu(N, R, [s2])
    assert np.array_equal(pg, pg2)                   # bytes past 2*N*repeats have no effect


@pytest.mark.parametrize("N,nbytes,R", [(512, 114688, 100), (4096, 16384 * 16, 30), (64, 4096, 32)])
def test_short_push_from_page_locked_memory_is_read_in_place(N, nbytes, R):
    """A short push that completes the integration out of page-locked memory (the reference-shaped Datastore page-locks
    its Buffers; config 1 is one such read) is read by the kernel in place, without the bounce copy: same bits as the
    same bytes pushed from pageable memory, the buffer may be overwritten as soon as rpf_push has returned, and a push
    that does NOT complete the integration still takes the bounce copy (so the caller is not held up)."""
    import ctypes as C
    ptr = C.c_void_p()
    assert pkg.lib().rpf_host_alloc(C.byref(ptr), nbytes) == 0
    try:
        pkg.lib().rpf_synth_fill_host(ptr, nbytes, 0, 0x5EED + 1, 1)
        pinned = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(nbytes,))
        pageable = pinned.copy()
        with Context(N) as c:
            c.reset(R); c.push(pageable); p0, d0 = c.finish()
            c.reset(R); c.push_ptr(ptr.value, nbytes); pinned[:] = 0; p1, d1 = c.finish()     # overwritten right after the push
            pinned[:] = pageable
            half = (R // 2) * 2 * N
            c.reset(R); c.push_ptr(ptr.value, half); c.push_ptr(ptr.value + half, nbytes - half); p2, d2 = c.finish()
        assert d0 == d1 == d2 == R
        assert np.array_equal(p0, p1)
        assert relmax(p0, p2) < 1e-6                  # (two launches: the float32 first-level sums are grouped differently)
        po, _ = ol.oracle_run_stream(N, R, pageable, buf_length=nbytes)
        assert perbin(p0, po) < TOL
    finally:
        pkg.lib().rpf_host_free(ptr)


@pytest.mark.parametrize("N,R", [(4096, 1), (4096, 2), (4096, 3), (4096, 295), (4096, 297), (4096, 593), (8192, 1),
                                 (8192, 149), (65536, 1), (65536, 9), (65536, 10), (65536, 19)])
def test_tail_shapes_of_the_pipelined_kernel(N, R):
    """Spectrum counts that leave stream B of the last pair empty, give some CTAs nothing or one pair, or sit one
    past a multiple of the grid: every ring slot is still filled (clamped re-reads) and nothing extra is
    accumulated.  One launch per case (a single push of R spectra)."""
    s = stream(2 * N * R, seed=R + N)
    pg, done = run_gpu(N, R, [s])
    po, done_o = ol.oracle_run_stream(N, R, s, buf_length=s.size)
    assert done == done_o == R
    assert relmax(pg, po) < TOL
    if R >= 9:                      # a single spectrum has bins far below the mean: per-bin bound only when averaged
        assert perbin(pg, po) < TOL
    else:
        assert perbin(pg, po) < 1e-3


@pytest.mark.parametrize("N,chunk", [(65536, 1638400), (4096, 16384 * 5 + 2), (512, 1000), (64, 6), (65536, 4096)])
def test_straddling_buffers_equal_contiguous(N, chunk):
    """Spectrum boundaries are independent of buffer boundaries (src/datastore.cxx:52,68)."""
    R = 5 if N == 65536 else 37
    s = stream(2 * N * R + 2 * (N // 2), seed=5)     # ends with half a spectrum
    chunks = [s[i:i + chunk] for i in range(0, s.size, chunk)]
    with Context(N) as c:
        c.reset(10 ** 9)
        for ch in chunks:
            c.push(ch)
        pg, done = c.finish()
        assert done == R and c.fft_pointer == N // 2
    po, _ = ol.oracle_run_stream(N, R, s, buf_length=s.size)
    assert perbin(pg, po) < TOL


def test_repeats_limit_and_reset():
    N = 4096
    s = stream(2 * N * 20, seed=2)
    with Context(N) as c:
        c.reset(7)
        c.push(s[: 2 * N * 5 + 100])
        assert c.repeats_done == 5
        c.push(s[2 * N * 5 + 100:])
        assert c.repeats_done == 7 and c.fft_pointer == 0
        c.push(s)                                     # discarded
        p7, done = c.finish()
        assert done == 7
        po, _ = ol.oracle_run_stream(N, 7, s, buf_length=s.size)
        assert perbin(p7, po) < TOL
        c.reset(3)
        p0, d0 = c.finish()
        assert d0 == 0 and np.all(p0 == 0)
        c.push(s)
        p3, d3 = c.finish()
        po3, _ = ol.oracle_run_stream(N, 3, s, buf_length=s.size)
        assert d3 == 3 and perbin(p3, po3) < TOL


def test_argument_errors():
    with pytest.raises(RpfError) as e:
        Context(511)
    assert e.value.code == -1
    with pytest.raises(RpfError) as e:
        Context((1 << 22) + 2)
    assert e.value.code == -2                         # valid for the reference, beyond the CUDA path's limit: loud
    with Context(64) as c:
        with pytest.raises(RpfError):
            c.push(np.zeros(3, np.uint8))
        with pytest.raises(RpfError):
            c.reset(-1)


def test_device_source_and_generator_identity():
    import torch
    N, R = 4096, 64
    nbytes = 2 * N * R
    dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    for mode in (0, 1, 2):
        pkg.synth_fill_device(dev.data_ptr(), nbytes, 0, seed=77, mode=mode)
        torch.cuda.synchronize()
        host = pkg.synth_fill_host(nbytes, 0, seed=77, mode=mode)
        assert np.array_equal(dev.cpu().numpy(), host)
    # unaligned offset / odd length
    pkg.synth_fill_device(dev.data_ptr() + 3, 1001, 12345, seed=9, mode=1)
    torch.cuda.synchronize()
    assert np.array_equal(dev.cpu().numpy()[3:1004], pkg.synth_fill_host(1001, 12345, seed=9, mode=1))
    pkg.synth_fill_device(dev.data_ptr(), nbytes, 0, seed=77, mode=0)
    torch.cuda.synchronize()
    host = pkg.synth_fill_host(nbytes, 0, seed=77, mode=0)
    with Context(N) as c:
        c.reset(R)
        c.push_device(dev.data_ptr(), nbytes)         # zero-copy path
        pd, done = c.finish()
        assert done == R
        c.reset(R)
        c.push_device(dev.data_ptr(), 2 * N * 3 + 10)  # leaves a carry -> staged path next
        c.push_device(dev.data_ptr() + 2 * N * 3 + 10, nbytes - (2 * N * 3 + 10))
        pd2, done2 = c.finish()
    po, _ = ol.oracle_run_stream(N, R, host, buf_length=nbytes)
    assert done2 == R and perbin(pd, po) < TOL and perbin(pd2, po) < TOL


def test_linearity_large_config2_shape():
    """Config 2 shape (4096 bins, 1e6 spectra) via a periodic device-resident ring:
    pwr(R*P spectra) == R * pwr(P spectra) (SURVEY section 4 'linearity in repeats')."""
    import torch
    N, P, reps = 4096, 1000, 1000                    # 1e6 spectra
    nbytes = 2 * N * P
    dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    pkg.synth_fill_device(dev.data_ptr(), nbytes, 0, seed=0x5EED + 2, mode=1)
    torch.cuda.synchronize()
    with Context(N) as c:
        c.reset(P)
        c.push_device(dev.data_ptr(), nbytes)
        p1, d1 = c.finish()
        c.reset(P * reps)
        for _ in range(reps):
            c.push_device(dev.data_ptr(), nbytes)
        pr, dr = c.finish()
    assert d1 == P and dr == P * reps
    assert np.max(np.abs(pr - reps * p1) / (reps * p1)) < 1e-9   # float64 second level
    host = pkg.synth_fill_host(2 * N * 50, 0, seed=0x5EED + 2, mode=1)
    po, _ = ol.oracle_run_stream(N, 50, host, buf_length=host.size)
    with Context(N) as c:
        c.reset(50)
        c.push(host)
        p50, _ = c.finish()
    assert perbin(p50, po) < TOL


@pytest.mark.parametrize("N,P,copies", [(65536, 64, 33), (8192, 128, 40), (16384, 64, 40)])
def test_fourstep_multibatch_single_launch(N, P, copies):
    """The four-step path loops over scratch-sized batches inside ONE rpf_push_device (csrc/rpf_capi.cu,
    launch_spectra; with a 256 MiB scratch: 512 spectra per batch at 65536 bins, 4096 at 8192): 2112 / 5120 / 2560 spectra in one call
    cross 4 / 1 / 1 batch boundaries.  The stream is `copies` repetitions of a P-spectra block, so
    pwr(copies * P spectra) == copies * pwr(P spectra); the P-block itself is pinned to the float64 oracle.
    The equality is not bit-exact here (unlike the repeated-push test below): one big launch partitions the
    spectra differently over the CTAs, so the fp32 first-level sums differ at the 1e-7 level."""
    import torch
    block = pkg.synth_fill_host(2 * N * P, 0, seed=0x5EED + 4, mode=1)
    dev = torch.from_numpy(block).cuda().repeat(copies)
    torch.cuda.synchronize()
    with Context(N, scratch_mib=256) as c:          # (default 1 GiB; 256 MiB keeps the multi-batch case small)
        c.reset(P)
        c.push_device(dev.data_ptr(), 2 * N * P)    # the scratch is sized for P spectra here and regrown below
        p1, d1 = c.finish()
        c.reset(P * copies)
        l0 = c.kernel_launches
        c.push_device(dev.data_ptr(), dev.numel())
        batches = (c.kernel_launches - l0) // 2
        pc, dc = c.finish()
    assert d1 == P and dc == P * copies
    assert batches == -(-P * copies // ((256 << 20) // (8 * N)))        # really went through the batch loop
    assert batches >= 2
    assert np.max(np.abs(pc - copies * p1) / (copies * p1)) < 2e-6
    po, _ = ol.oracle_run_stream(N, P, block, buf_length=block.size)
    assert perbin(p1, po) < TOL and perbin(pc / copies, po) < TOL
    assert np.argmax(pc) == np.argmax(po)


def test_window_linearity_config3_shape():
    """Config 3 shape (4096 bins, Hann window, 1e7 spectra): a 10,000-spectra device-resident block pushed
    1,000 times; float64 second level => exactly 1000 x the single-block result (to 1e-9), and the block's
    first 50 spectra are pinned to the oracle with the same window."""
    import torch
    N, P, reps = 4096, 10000, 1000
    w = (0.5 - 0.5 * np.cos(2 * np.pi * np.arange(N) / N)).astype(np.float32)
    nbytes = 2 * N * P
    dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    pkg.synth_fill_device(dev.data_ptr(), nbytes, 0, seed=0x5EED + 3, mode=1)
    torch.cuda.synchronize()
    with Context(N, window=w) as c:
        c.reset(P)
        c.push_device(dev.data_ptr(), nbytes)
        p1, d1 = c.finish()
        c.reset(P * reps)
        for _ in range(reps):
            c.push_device(dev.data_ptr(), nbytes)
        pr, dr = c.finish()
        c.reset(50)
        c.push_device(dev.data_ptr(), nbytes)            # truncated at 50 spectra by the repeats limit
        p50, d50 = c.finish()
    assert d1 == P and dr == P * reps and d50 == 50
    assert np.max(np.abs(pr - reps * p1) / (reps * p1)) < 1e-9
    host = pkg.synth_fill_host(2 * N * 50, 0, seed=0x5EED + 3, mode=1)
    po, _ = ol.oracle_run_stream(N, 50, host, window=w, buf_length=host.size)
    assert relmax(p50, po) < TOL and np.max(np.abs(p50 - po) / po) < 1e-4


@pytest.mark.parametrize("R", [1, 2, 3, 24, 295, 297, 593, 3000])
def test_tc_pass1_variant_matches_oracle_and_fp32_path(R):
    """RPF_FLAG_TC_PASS1 (first radix-16 pass on tcgen05 tensor cores, fp16 hi+lo DFT-16 matrix): same gate as the
    FP32 kernel -- 1e-5 against the float64 oracle -- on every 4096-bin shape the FP32 kernel is tested on."""
    N = 4096
    s = stream(2 * N * R, seed=R + 17, mode=R % 3)
    with Context(N, flags=pkg.capi.RPF_FLAG_TC_PASS1) as c:
        assert c.plan_name == "pipe4096_tc"
        c.reset(R)
        c.push(s)
        pt, done = c.finish()
    pf, _ = run_gpu(N, R, [s])
    assert done == R
    Ro = min(R, 600)
    if Ro == R:
        po, _ = ol.oracle_run_stream(N, R, s, buf_length=s.size)
        assert relmax(pt, po) < TOL
        if R >= 24 and R % 3 != 2:
            assert perbin(pt, po) < TOL
        assert np.argmax(pt) == np.argmax(po)
    assert relmax(pt, pf) < 2e-6                           # and against the FP32 kernel on the same bytes


def test_tc_pass1_variant_kats_and_stream_semantics():
    N = 4096
    F = pkg.capi.RPF_FLAG_TC_PASS1
    z = np.full(2 * N * 2, 127, np.uint8)
    with Context(N, flags=F) as c:
        c.reset(2)
        c.push(z)
        p, done = c.finish()
        assert done == 2 and np.all(p == 0.0)              # exact
        d = z[: 2 * N].copy()
        d[0::2] = 127 + 40
        c.reset(1)
        c.push(d)
        p, _ = c.finish()
        assert np.argmax(p) == N // 2 and p[N // 2] == pytest.approx((40 * N) ** 2, rel=1e-6)
        assert np.max(np.delete(p, N // 2)) <= 1e-9 * p[N // 2]
        e = z[: 2 * N].copy()
        e[0::2] = 0
        e[1::2] = 255                                      # extreme bytes: -127 / +128 on the two rails
        c.reset(1)
        c.push(e)
        p, _ = c.finish()
        assert np.argmax(p) == N // 2 and p[N // 2] == pytest.approx(N * N * (127.0 ** 2 + 128.0 ** 2), rel=1e-6)
        # straddling pushes + Parseval + repeats limit
        R = 37
        s = stream(2 * N * R + N, seed=5)
        c.reset(30)
        for off in range(0, s.size, 16384 * 5 + 2):
            c.push(s[off:off + 16384 * 5 + 2])
        p, done = c.finish()
        assert done == 30
        x = s[: 2 * N * 30].astype(np.int64) - 127
        assert p.sum() == pytest.approx(N * float((x * x).sum()), rel=1e-6)
        po, _ = ol.oracle_run_stream(N, 30, s, buf_length=s.size)
        assert perbin(p, po) < TOL
    w = np.ones(N, np.float32)
    with Context(N, window=w, flags=F) as c:
        assert c.plan_name == "pipe4096"                   # with a window the flag is ignored (FP32 kernel)


def test_write_data_matches_oracle():
    import ctypes as C
    N, R = 4096, 9
    s = stream(2 * N * R, seed=21)
    base = np.linspace(-1, 1, N)
    with Context(N) as c:
        c.reset(R)
        c.push(s)
        pwr, done = c.finish()
        for linear in (False, True):
            f, v = c.write_data(1420405752.0, 2000000, linear=linear, baseline=base)
            fo, vo = np.zeros(N), np.zeros(N)
            p = pwr.copy()
            dp = C.POINTER(C.c_double)
            ol.oracle().rpf_oracle_write_data(N, p.ctypes.data_as(dp), done, 1420405752.0, 2000000, int(linear),
                                              base.ctypes.data_as(dp), fo.ctypes.data_as(dp), vo.ctypes.data_as(dp))
            assert np.array_equal(f, fo)
            assert np.allclose(v, vo, rtol=1e-13, atol=1e-13)
        pwr2, _ = c.finish()
        assert np.array_equal(pwr, pwr2)               # write_data leaves the accumulator untouched


@pytest.mark.parametrize("N", [512, 1024, 2048])
@pytest.mark.parametrize("R", [1, 2, 3, 7, 8, 9, 15, 16, 17, 33, 100, 1185, 2371, 4737])
def test_pipelined_small_bins_tail_shapes(N, R):
    """N = 512 / 1024 / 2048 run the two-stream pipeline over group iterations of G = 8 / 4 / 2 spectra: counts that
    leave the last group partial, stream B of the last pair empty, some CTAs with nothing or a single pair, or sit one
    past a multiple of the grid.  Against the oracle, and against the round-1 lock-step kernel (RPF_FLAG_LOCKSTEP)."""
    s = stream(2 * N * R, seed=R * 7 + N)
    pg, done = run_gpu(N, R, [s], flags=pkg.capi.RPF_FLAG_PIPELINE)     # (forced: the default differs per N)
    with Context(N, flags=pkg.capi.RPF_FLAG_LOCKSTEP) as c:
        c.reset(R)
        c.push(s)
        pl, _ = c.finish()
    Ro = min(R, 300)
    if Ro == R:
        po, done_o = ol.oracle_run_stream(N, R, s, buf_length=s.size)
        assert done == done_o == R
        assert relmax(pg, po) < TOL
        assert perbin(pg, po) < (TOL if R >= 9 else 1e-3)
    assert done == R and relmax(pg, pl) < 2e-6
    x = s.astype(np.int64) - 127
    assert pg.sum() == pytest.approx(N * float((x * x).sum()), rel=1e-6)      # nothing lost, nothing counted twice


def test_matrix_row_f32_and_smallest_N():
    """The float32 matrix-mode row (src/acquisition.cxx:400-405: fpwrdb = pwrdb) and write_data at N = 2, where
    the reference's DC interpolation reads pwr[N/2 + 1] one past the end (the index wraps here)."""
    import ctypes as C
    N, R = 1024, 17
    s = stream(2 * N * R, seed=23)
    base = np.linspace(0.5, -0.5, N)
    with Context(N) as c:
        c.reset(R)
        c.push(s)
        pwr, done = c.finish()
        for linear in (False, True):
            row = c.write_row_f32(2400000, linear=linear, baseline=base)
            _, v = c.write_data(0.0, 2400000, linear=linear, baseline=base)
            fo, vo = np.zeros(N), np.zeros(N)
            p = pwr.copy()
            dp = C.POINTER(C.c_double)
            ol.oracle().rpf_oracle_write_data(N, p.ctypes.data_as(dp), done, 0.0, 2400000, int(linear),
                                              base.ctypes.data_as(dp), fo.ctypes.data_as(dp), vo.ctypes.data_as(dp))
            assert row.dtype == np.float32 and np.array_equal(row, v.astype(np.float32))
            assert np.allclose(row, vo.astype(np.float32), rtol=2e-6, atol=2e-6)
    with Context(2) as c:
        t = np.array([130, 120, 100, 140] * 5, np.uint8)
        c.reset(5)
        c.push(t)
        p2, d2 = c.finish()
        f, v = c.write_data(1.0e6, 1000, linear=True)
        assert d2 == 5 and np.all(np.isfinite(v))
        assert v[1] == pytest.approx(p2[0] / d2 / 2 / 1000, rel=1e-12)       # mean of the two neighbours = pwr[0]
        assert v[0] == pytest.approx(p2[0] / d2 / 2 / 1000, rel=1e-12)


@pytest.mark.parametrize("R", [1, 2, 3, 7, 8, 9, 15, 16, 17, 31, 33, 100, 257])
def test_cluster65536_matches_two_kernel_path_and_oracle(R):
    """RPF_FLAG_CLUSTER65536: 65536 bins in one launch of 16-CTA clusters (outer radix-16 handed over through distributed
    shared memory, csrc/cluster65536.cu; an experiment, not the default).  Spectrum counts that leave clusters idle,
    stream B of the last pair empty, or the clusters unevenly loaded: against the default two-kernel four-step path (both
    run the same arithmetic, only the float32 chunking of the accumulation differs), the oracle for small counts,
    Parseval."""
    N = 65536
    s = stream(2 * N * R, seed=R * 13 + 5, mode=1)
    pc, done = run_gpu(N, R, [s], flags=pkg.capi.RPF_FLAG_CLUSTER65536)
    pt, done_t = run_gpu(N, R, [s], flags=pkg.capi.RPF_FLAG_FOURSTEP)
    assert done == done_t == R
    assert relmax(pc, pt) < 2e-6 and perbin(pc, pt) < 2e-5
    if R <= 9:
        po, _ = ol.oracle_run_stream(N, R, s, buf_length=s.size)
        assert relmax(pc, po) < TOL
        assert perbin(pc, po) < (TOL if R >= 9 else 1e-3)
    x = s.astype(np.int64) - 127
    assert pc.sum() == pytest.approx(N * float((x * x).sum()), rel=1e-6)


@pytest.mark.parametrize("R", [1, 2, 3, 7, 9, 147, 148, 149, 297, 600, 1500])
def test_solo65536_matches_two_kernel_path_and_oracle(R):
    """RPF_FLAG_SOLO65536: 65536 bins in one launch, one CTA per spectrum (csrc/solo65536.cu): the outer radix-16 handed
    to the 16 inner transforms thread-privately through an in-place scratch whose layout alternates between [ko][a] and
    [a][ko], accumulators in tensor memory.  Spectrum counts below, at and above the CTA count (one, two, three ... and
    unequal numbers of spectra per CTA, i.e. both layouts and the re-read of the last spectrum): against the two-kernel
    four-step path, the oracle for small counts, Parseval."""
    N = 65536
    s = stream(2 * N * R, seed=R * 17 + 3, mode=1)
    with Context(N, flags=pkg.capi.RPF_FLAG_SOLO65536) as c:
        assert c.plan_name == "solo65536"
    ps, done = run_gpu(N, R, [s], flags=pkg.capi.RPF_FLAG_SOLO65536)
    pt, done_t = run_gpu(N, R, [s], flags=pkg.capi.RPF_FLAG_FOURSTEP)
    assert done == done_t == R
    # (per bin: with a handful of spectra some bins are orders of magnitude below the mean)
    assert relmax(ps, pt) < 2e-6 and perbin(ps, pt) < (2e-5 if R >= 9 else 1e-3)
    if R <= 9:
        po, _ = ol.oracle_run_stream(N, R, s, buf_length=s.size)
        assert relmax(ps, po) < TOL
        assert perbin(ps, po) < (TOL if R >= 9 else 1e-3)
    x = s.astype(np.int64) - 127
    assert ps.sum() == pytest.approx(N * float((x * x).sum()), rel=1e-6)


@pytest.mark.parametrize("R", [3, 149, 300])
def test_solo65536_window(R):
    """The same kernel with a Hann window (multiplication fused into the first radix-4 layer of the outer radix-16)."""
    N = 65536
    w = (0.5 - 0.5 * np.cos(2 * np.pi * np.arange(N) / N)).astype(np.float32)
    s = stream(2 * N * R, seed=R + 71, mode=1)
    ps, done = run_gpu(N, R, [s], window=w, flags=pkg.capi.RPF_FLAG_SOLO65536)
    pt, done_t = run_gpu(N, R, [s], window=w, flags=pkg.capi.RPF_FLAG_FOURSTEP)
    assert done == done_t == R
    # (per bin: with a handful of spectra some bins are orders of magnitude below the mean)
    assert relmax(ps, pt) < 2e-6 and perbin(ps, pt) < (2e-5 if R >= 9 else 1e-3)
    if R <= 3:
        po, _ = ol.oracle_run_stream(N, R, s, window=w, buf_length=s.size)
        assert relmax(ps, po) < TOL


@pytest.mark.parametrize("N", [8192, 16384, 32768])
@pytest.mark.parametrize("R", [1, 3, 9, 148, 149, 300, 1000])
@pytest.mark.parametrize("windowed", [False, True])
def test_solo_kernel_other_outer_radices(N, R, windowed):
    """The one-CTA-per-spectrum kernel with outer radix 2, 4, 8 (an outer stage = 8 / 4 / 2 radix-M butterflies, 1 / 2 / 4
    iterations per spectrum, in-place scratch of M x 16 blocks with the two layouts of Geo<M>): against the two-kernel
    path, the oracle for small counts, Parseval."""
    if windowed and R not in (3, 149, 300):
        pytest.skip("window: three counts")
    w = (0.5 - 0.5 * np.cos(2 * np.pi * np.arange(N) / N)).astype(np.float32) if windowed else None
    s = stream(2 * N * R, seed=R * 19 + N, mode=1)
    with Context(N, flags=pkg.capi.RPF_FLAG_SOLO65536) as c:
        assert c.plan_name == "solo"
    ps, done = run_gpu(N, R, [s], window=w, flags=pkg.capi.RPF_FLAG_SOLO65536)
    pt, done_t = run_gpu(N, R, [s], window=w, flags=pkg.capi.RPF_FLAG_FOURSTEP)
    assert done == done_t == R
    assert relmax(ps, pt) < 2e-6 and perbin(ps, pt) < (2e-5 if R >= 9 else 1e-3)
    if R <= 9:
        po, _ = ol.oracle_run_stream(N, R, s, window=w, buf_length=s.size)
        assert relmax(ps, po) < TOL
        assert perbin(ps, po) < (TOL if R >= 9 else 1e-3)
    if not windowed:
        x = s.astype(np.int64) - 127
        assert ps.sum() == pytest.approx(N * float((x * x).sum()), rel=1e-6)


@pytest.mark.parametrize("N", [8192, 32768, 65536])
def test_solo_kernel_long_launch_flushes_tensor_memory(N):
    """A launch long enough for every CTA to flush its tensor-memory accumulators inside the loop (every 128 spectra
    per CTA): 148 x 128 + 3000 spectra, device-resident.  Exact Parseval over the whole stream, the two-kernel path on
    the same bytes, linearity (the stream is a block repeated, so pwr == copies x pwr(block) to float32-grouping
    accuracy)."""
    import torch
    P = 148 * 8 + 17                                   # block: every CTA gets 8 or 9 spectra
    copies = 17                                        # 20,417 spectra: 137-138 per CTA
    block = pkg.synth_fill_host(2 * N * P, 0, seed=N + 77, mode=1)
    dev = torch.from_numpy(block).cuda().repeat(copies)
    torch.cuda.synchronize()
    with Context(N, flags=pkg.capi.RPF_FLAG_SOLO65536) as c:
        c.reset(P); c.push_device(dev.data_ptr(), 2 * N * P); p1, d1 = c.finish()
        c.reset(P * copies); c.push_device(dev.data_ptr(), dev.numel()); pc, dc = c.finish()
    with Context(N, flags=pkg.capi.RPF_FLAG_FOURSTEP) as c:
        c.reset(P * copies); c.push_device(dev.data_ptr(), dev.numel()); pt, dt = c.finish()
    assert d1 == P and dc == dt == P * copies
    x = block.astype(np.int64) - 127
    assert pc.sum() == pytest.approx(copies * N * float((x * x).sum()), rel=1e-6)
    assert relmax(pc, copies * p1) < 2e-6 and perbin(pc, copies * p1) < 2e-5
    assert relmax(pc, pt) < 2e-6 and perbin(pc, pt) < 2e-5


def test_solo65536_kats_pushes_and_reuse():
    """The one-CTA-per-spectrum kernel on known answers (zero input, DC, a tone: bin exact), over several pushes that
    straddle spectra, and reused after rpf_reset (tensor-memory accumulators and float64 rows start from zero)."""
    N, F = 65536, pkg.capi.RPF_FLAG_SOLO65536
    with Context(N, flags=F) as c:
        z = np.full(2 * N * 3, 127, dtype=np.uint8)
        c.reset(3); c.push(z); p, done = c.finish()
        assert done == 3 and np.all(p == 0.0)
        d = np.full(2 * N * 2, 127, dtype=np.uint8); d[0::2] = 137
        c.reset(2); c.push(d); p, done = c.finish()
        assert done == 2 and int(np.argmax(p)) == N // 2 and p[N // 2] == pytest.approx(2 * (10.0 * N) ** 2, rel=1e-6)
        assert np.all(np.delete(p, N // 2) <= 1e-6 * p[N // 2])
        k0 = 12345
        n = np.arange(N)
        t = np.empty(2 * N, dtype=np.uint8)
        t[0::2] = np.round(127 + 100 * np.cos(2 * np.pi * k0 * n / N)).astype(np.uint8)
        t[1::2] = np.round(127 + 100 * np.sin(2 * np.pi * k0 * n / N)).astype(np.uint8)
        c.reset(1); c.push(t); p, done = c.finish()
        assert done == 1 and int(np.argmax(p)) == (k0 + N // 2) % N
        R = 301
        s = stream(2 * N * R, seed=99, mode=1)
        cuts = [0, 1000, 2 * N + 6, 2 * N * 150 + 2, 2 * N * 150 + 4, s.size]
        c.reset(R)
        for a, b in zip(cuts[:-1], cuts[1:]):
            c.push(s[a:b])
        p1, done = c.finish()
        assert done == R
        c.reset(R); c.push(s); p2, done = c.finish()
        assert done == R
    pt, _ = run_gpu(N, R, [s], flags=pkg.capi.RPF_FLAG_FOURSTEP)
    assert relmax(p1, pt) < 2e-6 and perbin(p1, pt) < 2e-5
    assert relmax(p1, p2) < 1e-6


def test_65536_default_mixes_both_paths():
    """Without a flag a 65536-bin context sends launches of at least 96 spectra to the one-CTA-per-spectrum kernel and
    shorter ones (the reference's 1,638,400-byte reads hold 12.5 spectra) to the two-kernel path; both add into the same
    private rows.  One integration fed by pushes of both kinds must equal either path alone and the oracle on a prefix."""
    N, R = 65536, 12 + 300 + 7 + 150
    s = stream(2 * N * R, seed=4242, mode=1)
    cuts = [0, 2 * N * 12, 2 * N * 312 + 1000, 2 * N * 319 + 1000, s.size]
    with Context(N) as c:
        assert c.plan_name == "solo65536"
        c.reset(R)
        l0 = c.kernel_launches
        for a, b in zip(cuts[:-1], cuts[1:]):
            c.push(s[a:b])
        launches = c.kernel_launches - l0
        pm, done = c.finish()
        assert done == R and launches == 2 + 1 + 2 + 1 + 1      # two-kernel, solo, two-kernel, solo; + the reduction, enqueued by the push that completes the integration
        c.reset(8)
        c.push(s[:2 * N * 8])
        p8, _ = c.finish()
    ps, _ = run_gpu(N, R, [s], flags=pkg.capi.RPF_FLAG_SOLO65536)
    pt, _ = run_gpu(N, R, [s], flags=pkg.capi.RPF_FLAG_FOURSTEP)
    assert relmax(pm, ps) < 2e-6 and perbin(pm, ps) < 2e-5
    assert relmax(pm, pt) < 2e-6 and perbin(pm, pt) < 2e-5
    po, _ = ol.oracle_run_stream(N, 8, s[:2 * N * 8], buf_length=2 * N * 8)
    assert relmax(p8, po) < TOL


@pytest.mark.parametrize("N,R", [(8192, 37), (16384, 300), (32768, 5), (65536, 1), (65536, 19), (65536, 700)])
@pytest.mark.parametrize("windowed", [False, True])
def test_tma_outer_pass_is_bit_identical_to_the_plain_one(N, R, windowed):
    """The four-step outer pass as a TMA-staged streaming kernel (csrc/outer_tma.cu) performs the arithmetic of the
    round-1 kernel (RPF_FLAG_LOCKSTEP selects it) in the same order: the spectra must agree bit for bit, for every
    outer radix, with a window, for counts that leave column ranges without work or cross a scratch batch."""
    w = (0.5 - 0.5 * np.cos(2 * np.pi * np.arange(N) / N)).astype(np.float32) if windowed else None
    s = stream(2 * N * R, seed=N + R, mode=1)
    F = pkg.capi.RPF_FLAG_FOURSTEP                    # (65536 bins: keep both runs on the two-kernel path)
    pn, done = run_gpu(N, R, [s], window=w, flags=F)
    po, done_o = run_gpu(N, R, [s], window=w, flags=F | pkg.capi.RPF_FLAG_LOCKSTEP)
    assert done == done_o == R
    assert np.array_equal(pn, po)
    if R <= 5:
        pr, _ = ol.oracle_run_stream(N, R, s, window=w, buf_length=s.size)
        assert relmax(pn, pr) < TOL


@pytest.mark.parametrize("N", [16, 32, 64, 128, 256, 512, 1024, 2048])
@pytest.mark.parametrize("R", [1, 2, 3, 7, 8, 9, 15, 16, 17, 33, 100, 1185, 2371, 4737, 75777])
def test_batched_small_bins_on_the_4096_loop(N, R):
    """RPF_FLAG_BATCHED4096 (the default from 16 to 256 and at 1024 / 2048 bins): N bins run on the 4096-bin kernel's loop
    with 4096 / N consecutive spectra per ring slot (N <= 256: pass 1 is empty; N < 256: pass 2 is 16 / R2 radix-R2
    butterflies).  Spectrum counts that leave the stream's last slot partial (only some of its spectra exist), stream B
    of the last pair empty, CTAs with nothing or a single pair: against the oracle, the lock-step kernel, Parseval."""
    F = pkg.capi.RPF_FLAG_BATCHED4096
    s = stream(2 * N * R, seed=R * 11 + N)
    with Context(N, flags=F) as c:
        assert c.plan_name == "pipe4096_batched"
        c.reset(R)
        c.push(s)
        pg, done = c.finish()
    pl, _ = run_gpu(N, R, [s], flags=pkg.capi.RPF_FLAG_LOCKSTEP)
    if R <= 300:
        po, done_o = ol.oracle_run_stream(N, R, s, buf_length=s.size)
        assert done == done_o == R
        assert relmax(pg, po) < TOL
        assert perbin(pg, po) < (TOL if R >= 9 else 1e-3)
    assert done == R and relmax(pg, pl) < 2e-6
    x = s.astype(np.int64) - 127
    assert pg.sum() == pytest.approx(N * float((x * x).sum()), rel=1e-6)      # nothing lost, nothing counted twice


@pytest.mark.parametrize("N", [16, 32, 64, 128, 256, 512, 1024, 2048])
def test_batched_small_bins_window_kats_and_straddling_pushes(N):
    """The same path with a Hann window, the analytic KATs (tone bin, zero input), and odd-sized pushes that straddle
    spectra and ring slots (the carry is resolved on the host, so every launch starts on a spectrum boundary -- but not
    on a slot boundary of the stream: slots are per launch)."""
    F = pkg.capi.RPF_FLAG_BATCHED4096
    w = (0.5 - 0.5 * np.cos(2 * np.pi * np.arange(N) / N)).astype(np.float32)
    R = 37
    s = stream(2 * N * R, seed=N + 3, mode=1)
    pg, done = run_gpu(N, R, [s], window=w, flags=F)
    po, _ = ol.oracle_run_stream(N, R, s, window=w, buf_length=s.size)
    assert done == R and relmax(pg, po) < TOL and perbin(pg, po) < TOL
    t = stream(2 * N * 12, seed=5, mode=2)                                   # tone at +fs/8
    pt, _ = run_gpu(N, 12, [t], flags=F)
    assert int(np.argmax(pt)) == N // 2 + N // 8
    z = np.full(2 * N * 5, 127, np.uint8)
    pz, _ = run_gpu(N, 5, [z], flags=F)
    assert np.all(pz == 0.0)
    chunks = [s[i:i + 10006] for i in range(0, s.size, 10006)]
    pc, done_c = run_gpu(N, R, chunks, flags=F)
    pw, _ = run_gpu(N, R, [s], flags=F)
    assert done_c == R and relmax(pc, pw) < 2e-6
