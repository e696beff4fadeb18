"""rtl-power-fftw_b200 -- B200-native FFT worker for rtl_power_fftw.

The product is the C-ABI shared library ``librpf_b200.so`` (include/rpf_b200.h),
hand-written sm_100a CUDA.  This module is only a ctypes view of that ABI for
tests and bench.py; the C++ host side that mirrors the reference's
Datastore/Acquisition lives in ``host/``.  The directory name is not a Python
identifier: ``importlib.import_module("rtl-power-fftw_b200")``.

There is no CPU fallback: if the library is missing, importing ``capi`` raises.
"""
from . import capi  # noqa: F401
from .capi import (RpfError, Context, lib, synth_fill_host, synth_fill_device, LIB_PATH,  # noqa: F401
                   EXPORTED_SYMBOLS, comm_unique_id)
