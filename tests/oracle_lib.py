"""ctypes access to the CPU checker under oracle/ (test infrastructure only).

Two libraries:
  * oracle/liboracle.so       -- rpf_oracle.c, the restatement of
                                 src/datastore.cxx:48-96 (+ producer arithmetic)
  * oracle/_ref/librpf_ref.so -- the reference's own src/datastore.cxx compiled
                                 against the fftw3.h shim (stand-in FFT)
Both are built by `make -C oracle` (done by __graft_entry__.build()).
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
_u8p = C.POINTER(C.c_uint8)
_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)
_i64p = C.POINTER(C.c_int64)


def _build():
    subprocess.run(["make", "-C", ORACLE_DIR], check=True, stdout=subprocess.DEVNULL)


def _load(path):
    if not os.path.exists(path):
        _build()
    return C.CDLL(path)


_oracle = None
_ref = None


def oracle():
    global _oracle
    if _oracle is None:
        lib = _load(os.path.join(ORACLE_DIR, "liboracle.so"))
        lib.rpf_oracle_create.restype = C.c_void_p
        lib.rpf_oracle_create.argtypes = [C.c_int, C.c_int64, _f32p, C.c_int]
        lib.rpf_oracle_destroy.argtypes = [C.c_void_p]
        lib.rpf_oracle_reset.argtypes = [C.c_void_p, C.c_int64]
        lib.rpf_oracle_push.argtypes = [C.c_void_p, _u8p, C.c_size_t]
        lib.rpf_oracle_push.restype = C.c_int
        lib.rpf_oracle_pwr.restype = _f64p
        lib.rpf_oracle_pwr.argtypes = [C.c_void_p]
        lib.rpf_oracle_repeats_done.restype = C.c_int64
        lib.rpf_oracle_repeats_done.argtypes = [C.c_void_p]
        lib.rpf_oracle_fft_pointer.restype = C.c_int
        lib.rpf_oracle_fft_pointer.argtypes = [C.c_void_p]
        lib.rpf_oracle_dft.argtypes = [C.c_int, _f64p, _f64p, C.c_int]
        lib.rpf_oracle_plan_buf_length.restype = C.c_int64
        lib.rpf_oracle_plan_buf_length.argtypes = [C.c_int, C.c_int64, C.c_int64, C.c_int]
        lib.rpf_oracle_read_sizes.restype = C.c_int64
        lib.rpf_oracle_read_sizes.argtypes = [C.c_int, C.c_int64, C.c_int64, _i64p, C.c_int64]
        lib.rpf_oracle_write_data.argtypes = [C.c_int, _f64p, C.c_int64, C.c_double, C.c_int, C.c_int,
                                              _f64p, _f64p, _f64p]
        lib.rpf_oracle_plan_repeats.restype = C.c_int64
        lib.rpf_oracle_plan_repeats.argtypes = [C.c_int, C.c_int, C.c_double]
        lib.rpf_oracle_plan_freqs.restype = C.c_int
        lib.rpf_oracle_plan_freqs.argtypes = [C.c_int64, C.c_int64, C.c_int, C.c_double, _i64p, C.c_int]
        _oracle = lib
    return _oracle


def have_ref():
    p = os.path.join(ORACLE_DIR, "_ref", "librpf_ref.so")
    if os.path.exists(p):
        return True
    if os.path.exists("/root/reference/src/datastore.cxx"):
        _build()
        return os.path.exists(p)
    return False


def ref():
    global _ref
    if _ref is None:
        lib = _load(os.path.join(ORACLE_DIR, "_ref", "librpf_ref.so"))
        lib.rpf_ref_run.restype = C.c_int
        lib.rpf_ref_run.argtypes = [C.c_int, C.c_int64, _f32p, C.c_int, C.c_int, _u8p, C.c_size_t, C.c_int,
                                    _f64p, _i64p, C.POINTER(C.c_int), _i64p]
        lib.rpf_ref_run_parallel.restype = C.c_int
        lib.rpf_ref_run_parallel.argtypes = [C.c_int, C.c_int, C.c_int64, _f32p, C.c_int, C.c_int, _u8p,
                                             C.c_size_t, _f64p, _i64p]
        _ref = lib
    return _ref


def _u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a, a.ctypes.data_as(_u8p)


def _win(window):
    if window is None:
        return None, None
    w = np.ascontiguousarray(window, dtype=np.float32)
    return w, w.ctypes.data_as(_f32p)


class OracleWorker:
    """Restated Datastore worker (oracle/rpf_oracle.c)."""

    def __init__(self, N, repeats, window=None, fft_kind=0):
        self.N = N
        self._w, wp = _win(window)
        self._h = oracle().rpf_oracle_create(N, repeats, wp, fft_kind)
        if not self._h:
            raise ValueError("oracle: invalid N")

    def reset(self, repeats):
        oracle().rpf_oracle_reset(self._h, repeats)

    def push(self, buf):
        a, p = _u8(buf)
        rc = oracle().rpf_oracle_push(self._h, p, a.size)
        if rc:
            raise ValueError("oracle push rc=%d" % rc)

    @property
    def pwr(self):
        return np.ctypeslib.as_array(oracle().rpf_oracle_pwr(self._h), shape=(self.N,)).copy()

    @property
    def repeats_done(self):
        return oracle().rpf_oracle_repeats_done(self._h)

    @property
    def fft_pointer(self):
        return oracle().rpf_oracle_fft_pointer(self._h)

    def close(self):
        if self._h:
            oracle().rpf_oracle_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()


def oracle_read_sizes(N, repeats, buf_length):
    n = oracle().rpf_oracle_read_sizes(N, repeats, buf_length, None, 0)
    out = np.zeros(n, dtype=np.int64)
    oracle().rpf_oracle_read_sizes(N, repeats, buf_length, out.ctypes.data_as(_i64p), n)
    return out


def oracle_plan_freqs(startfreq, stopfreq, samplerate, min_overlap):
    n = oracle().rpf_oracle_plan_freqs(startfreq, stopfreq, samplerate, min_overlap, None, 0)
    out = np.zeros(n, dtype=np.int64)
    oracle().rpf_oracle_plan_freqs(startfreq, stopfreq, samplerate, min_overlap, out.ctypes.data_as(_i64p), n)
    return [int(x) for x in out]


def oracle_run_stream(N, repeats, stream, window=None, buf_length=None, fft_kind=0, wrap=False):
    """Feed `stream` through the restated worker in reads sized as the reference's
    producer would size them (src/acquisition.cxx:273-302)."""
    stream = np.ascontiguousarray(stream, dtype=np.uint8)
    if buf_length is None:
        buf_length = oracle().rpf_oracle_plan_buf_length(N, repeats, 16384 * 100, 0)
    w = OracleWorker(N, repeats, window, fft_kind)
    off = 0
    for sz in oracle_read_sizes(N, repeats, buf_length):
        sz = int(sz)
        if wrap:
            idx = (np.arange(off, off + sz) % stream.size)
            chunk = stream[idx]
        else:
            chunk = np.full(sz, 127, dtype=np.uint8)
            avail = max(0, min(sz, stream.size - off))
            chunk[:avail] = stream[off:off + avail]
        w.push(chunk)
        off += sz
    pwr, done = w.pwr, w.repeats_done
    w.close()
    return pwr, done


def ref_run(N, repeats, stream, window=None, buffers=5, buf_length=None, wrap=False):
    """The reference's own Datastore::fftThread (oracle/_ref) over `stream`."""
    a, p = _u8(stream)
    _w, wp = _win(window)
    if buf_length is None:
        buf_length = oracle().rpf_oracle_plan_buf_length(N, repeats, 16384 * 100, 0)
    pwr = np.zeros(N, dtype=np.float64)
    done = C.c_int64(0)
    reads = C.c_int64(0)
    hist = (C.c_int * (buffers + 1))()
    rc = ref().rpf_ref_run(N, repeats, wp, buffers, int(buf_length), p, a.size, 1 if wrap else 0,
                           pwr.ctypes.data_as(_f64p), C.byref(done), hist, C.byref(reads))
    if rc:
        raise RuntimeError("rpf_ref_run rc=%d" % rc)
    return pwr, done.value, list(hist), reads.value


def numpy_power(N, stream, repeats, window=None):
    """Independent float64 statement of the path via numpy: fftshift(|FFT|^2)."""
    x = np.asarray(stream[: 2 * N * repeats], dtype=np.float64).reshape(repeats, N, 2)
    z = (x[..., 0] - 127.0) + 1j * (x[..., 1] - 127.0)
    if window is not None:
        # reference multiplies in float32 (src/datastore.cxx:76-77)
        w = np.asarray(window, dtype=np.float32)
        zr = (z.real.astype(np.float32) * w).astype(np.float64)
        zi = (z.imag.astype(np.float32) * w).astype(np.float64)
        z = zr + 1j * zi
    X = np.fft.fft(z, axis=1)
    return np.fft.fftshift((X.real ** 2 + X.imag ** 2).sum(axis=0))
