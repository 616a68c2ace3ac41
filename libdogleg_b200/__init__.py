"""libdogleg_b200 -- B200-native implementation of libdogleg's per-iteration hot path.

The product is the C-ABI shared library ``libdogleg_b200/libdogleg.so`` (sources in
``libdogleg_b200/csrc``): the unchanged ``dogleg.h`` API (reference dogleg.h:214-392)
plus the additive ``dogleg_gpu.h``. This package is only the thin ctypes binding the
tests and ``bench.py`` use; it contains no numerical code and no CPU fallback.
"""
from .ffi import (Parameters, Scalars, load, lib_path, build, default_parameters,
                  SOLVE_DENSE, SOLVE_SPARSE, SOLVE_DENSE_PRODUCTS, DEBUG_VNLOG)

__all__ = ["Parameters", "Scalars", "load", "lib_path", "build", "default_parameters",
           "SOLVE_DENSE", "SOLVE_SPARSE", "SOLVE_DENSE_PRODUCTS", "DEBUG_VNLOG"]
