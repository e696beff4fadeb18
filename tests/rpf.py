"""Import helper: the package directory name is not a Python identifier."""
import importlib

pkg = importlib.import_module("rtl-power-fftw_b200")
Context = pkg.Context
RpfError = pkg.RpfError
