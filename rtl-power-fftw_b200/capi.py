"""ctypes binding of include/rpf_b200.h (one-to-one; no logic of its own)."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "librpf_b200.so")
RPF_FLAG_TC_PASS1 = 1
RPF_FLAG_PIPELINED_PUSH = 2
RPF_FLAG_LOCKSTEP = 4
RPF_FLAG_PIPELINE = 8
RPF_FLAG_CLUSTER65536 = 16
RPF_FLAG_BATCHED4096 = 32
RPF_FLAG_SOLO65536 = 64
RPF_FLAG_FOURSTEP = 128
RPF_MAX_DEVICES = 16

EXPORTED_SYMBOLS = [
    "rpf_abi_version", "rpf_create", "rpf_destroy", "rpf_reset", "rpf_push", "rpf_push_device", "rpf_sync",
    "rpf_finish", "rpf_finish_device", "rpf_repeats_done", "rpf_fft_pointer", "rpf_set_stream", "rpf_strerror",
    "rpf_kernel_launches", "rpf_plan_name", "rpf_write_data", "rpf_synth_fill_host", "rpf_synth_fill_device",
    "rpf_host_alloc", "rpf_host_free", "rpf_host_register", "rpf_host_unregister",
    "rpf_comm_unique_id", "rpf_comm_init_rank", "rpf_device_count", "rpf_write_row_f32",
]

RPF_OK, RPF_ERR_INVALID_ARG, RPF_ERR_UNSUPPORTED_N, RPF_ERR_CUDA, RPF_ERR_NOMEM, RPF_ERR_STATE = 0, -1, -2, -3, -4, -5


class RpfConfig(C.Structure):
    _fields_ = [("N", C.c_int32), ("device", C.c_int32), ("window", C.POINTER(C.c_float)),
                ("buf_length", C.c_int64), ("flags", C.c_uint32), ("n_devices", C.c_int32),
                ("devices", C.c_int32 * RPF_MAX_DEVICES), ("scratch_mib", C.c_int32), ("reserved", C.c_int32 * 3)]


class RpfError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("rpf error %d: %s" % (code, msg))
        self.code = code


_lib = None


def lib():
    """Load librpf_b200.so (built in-tree by __graft_entry__.build()).  Fails loudly."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, u8p, f64p, i64p = C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64)
    L.rpf_abi_version.restype = C.c_int
    L.rpf_create.argtypes = [C.POINTER(RpfConfig), C.POINTER(vp)]
    L.rpf_destroy.argtypes = [vp]
    L.rpf_destroy.restype = None
    L.rpf_reset.argtypes = [vp, C.c_int64]
    L.rpf_push.argtypes = [vp, u8p, C.c_size_t]
    L.rpf_push_device.argtypes = [vp, u8p, C.c_size_t]
    L.rpf_sync.argtypes = [vp]
    L.rpf_finish.argtypes = [vp, f64p, i64p]
    L.rpf_finish_device.argtypes = [vp, f64p, i64p]
    L.rpf_repeats_done.argtypes = [vp]
    L.rpf_repeats_done.restype = C.c_int64
    L.rpf_fft_pointer.argtypes = [vp]
    L.rpf_fft_pointer.restype = C.c_int64
    L.rpf_set_stream.argtypes = [vp, vp]
    L.rpf_strerror.argtypes = [vp]
    L.rpf_strerror.restype = C.c_char_p
    L.rpf_kernel_launches.argtypes = [vp]
    L.rpf_kernel_launches.restype = C.c_int64
    L.rpf_plan_name.argtypes = [vp]
    L.rpf_plan_name.restype = C.c_char_p
    L.rpf_write_data.argtypes = [vp, C.c_double, C.c_int32, C.c_int32, f64p, f64p, f64p]
    L.rpf_synth_fill_host.argtypes = [u8p, C.c_size_t, C.c_uint64, C.c_uint64, C.c_int32]
    L.rpf_synth_fill_host.restype = None
    L.rpf_synth_fill_device.argtypes = [u8p, C.c_size_t, C.c_uint64, C.c_uint64, C.c_int32, C.c_int32, vp]
    L.rpf_host_alloc.argtypes = [C.POINTER(vp), C.c_size_t]
    L.rpf_host_free.argtypes = [vp]
    L.rpf_host_free.restype = None
    L.rpf_host_register.argtypes = [vp, C.c_size_t]
    L.rpf_host_unregister.argtypes = [vp]
    L.rpf_comm_unique_id.argtypes = [vp, C.c_size_t]
    L.rpf_comm_init_rank.argtypes = [vp, vp, C.c_size_t, C.c_int32, C.c_int32]
    L.rpf_device_count.argtypes = [vp]
    L.rpf_device_count.restype = C.c_int32
    L.rpf_write_row_f32.argtypes = [vp, C.c_int32, C.c_int32, f64p, vp]
    for name in ("rpf_create", "rpf_reset", "rpf_push", "rpf_push_device", "rpf_sync", "rpf_finish",
                 "rpf_finish_device", "rpf_set_stream", "rpf_write_data", "rpf_synth_fill_device",
                 "rpf_host_alloc", "rpf_host_register", "rpf_host_unregister", "rpf_comm_unique_id",
                 "rpf_comm_init_rank", "rpf_write_row_f32"):
        getattr(L, name).restype = C.c_int
    _lib = L
    return L


def synth_fill_host(nbytes, stream_offset=0, seed=0x5EED, mode=0):
    out = np.empty(nbytes, dtype=np.uint8)
    lib().rpf_synth_fill_host(out.ctypes.data, nbytes, stream_offset, seed, mode)
    return out


def synth_fill_device(ptr, nbytes, stream_offset=0, seed=0x5EED, mode=0, device=0, stream=0):
    rc = lib().rpf_synth_fill_device(ptr, nbytes, stream_offset, seed, mode, device, stream)
    if rc:
        raise RpfError(rc, "rpf_synth_fill_device")


def comm_unique_id():
    buf = (C.c_char * 128)()
    rc = lib().rpf_comm_unique_id(buf, 128)
    if rc:
        raise RpfError(rc, lib().rpf_strerror(None).decode())
    return bytes(buf)


class Context:
    """Thin RAII wrapper of rpf_ctx.  Method names follow the C ABI."""

    def __init__(self, N, window=None, device=0, buf_length=0, flags=0, devices=None, scratch_mib=0):
        self.N = N
        self._h = C.c_void_p()
        cfg = RpfConfig()
        cfg.N, cfg.device, cfg.buf_length, cfg.flags = N, device, buf_length, flags
        cfg.scratch_mib = scratch_mib
        if devices is not None:
            cfg.n_devices = len(devices)
            for i, d in enumerate(devices):
                cfg.devices[i] = d
        self._w = None
        if window is not None:
            self._w = np.ascontiguousarray(window, dtype=np.float32)
            if self._w.size != N:
                raise ValueError("window must have N values")
            cfg.window = self._w.ctypes.data_as(C.POINTER(C.c_float))
        rc = lib().rpf_create(C.byref(cfg), C.byref(self._h))
        if rc:
            raise RpfError(rc, lib().rpf_strerror(None).decode())

    def _chk(self, rc):
        if rc:
            raise RpfError(rc, lib().rpf_strerror(self._h).decode())

    def reset(self, repeats):
        self._chk(lib().rpf_reset(self._h, repeats))

    def push(self, buf):
        """buf: numpy uint8 array (host) -- rpf_push."""
        a = np.ascontiguousarray(buf, dtype=np.uint8)
        self._chk(lib().rpf_push(self._h, a.ctypes.data, a.size))

    def push_ptr(self, host_ptr, nbytes):
        self._chk(lib().rpf_push(self._h, host_ptr, nbytes))

    def push_device(self, dev_ptr, nbytes):
        self._chk(lib().rpf_push_device(self._h, dev_ptr, nbytes))

    def sync(self):
        self._chk(lib().rpf_sync(self._h))

    def finish(self):
        pwr = np.empty(self.N, dtype=np.float64)
        done = C.c_int64(0)
        self._chk(lib().rpf_finish(self._h, pwr.ctypes.data, C.byref(done)))
        return pwr, done.value

    def finish_device(self, dev_ptr):
        done = C.c_int64(0)
        self._chk(lib().rpf_finish_device(self._h, dev_ptr, C.byref(done)))
        return done.value

    def write_data(self, tuned_freq, samplerate, linear=False, baseline=None):
        f = np.empty(self.N)
        v = np.empty(self.N)
        b = None if baseline is None else np.ascontiguousarray(baseline, dtype=np.float64)
        self._chk(lib().rpf_write_data(self._h, tuned_freq, samplerate, int(linear),
                                       None if b is None else b.ctypes.data, f.ctypes.data, v.ctypes.data))
        return f, v

    def write_row_f32(self, samplerate, linear=False, baseline=None):
        row = np.empty(self.N, dtype=np.float32)
        b = None if baseline is None else np.ascontiguousarray(baseline, dtype=np.float64)
        self._chk(lib().rpf_write_row_f32(self._h, samplerate, int(linear), None if b is None else b.ctypes.data,
                                          row.ctypes.data))
        return row

    def comm_init_rank(self, unique_id, rank, nranks):
        """unique_id: 128 bytes from comm_unique_id() on rank 0 (shipped by the caller)."""
        buf = (C.c_char * 128).from_buffer_copy(bytes(unique_id))
        self._chk(lib().rpf_comm_init_rank(self._h, buf, 128, rank, nranks))

    @property
    def device_count(self):
        return lib().rpf_device_count(self._h)

    def set_stream(self, cuda_stream):
        self._chk(lib().rpf_set_stream(self._h, cuda_stream))

    @property
    def repeats_done(self):
        return lib().rpf_repeats_done(self._h)

    @property
    def fft_pointer(self):
        return lib().rpf_fft_pointer(self._h)

    @property
    def kernel_launches(self):
        return lib().rpf_kernel_launches(self._h)

    @property
    def plan_name(self):
        return lib().rpf_plan_name(self._h).decode()

    def close(self):
        if getattr(self, "_h", None):
            lib().rpf_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
