"""gcm_b200 — B200-native grid-characteristic time stepping (hot path of AlexanderKazakov/gcm).

The package is a thin ctypes binding over two native libraries built in-tree by `python -m gcm_b200.build`:
libgcm_b200.so (CUDA sm_100a kernels behind the C ABI of include/gcm_b200.h) and libgcm_b200_host.so (the
C++ mirror of the reference's Task/Engine interface).  Nothing is computed in Python and there is no CPU
fallback: `library()` raises ImportError when the native libraries are missing and `Context()` raises when
there is no CUDA device.
"""
from .capi import (Context, CubicBody, GcmError, HostEngine, Library, Q_PRESSURE_TRACE,  # noqa: F401
                   SimplexHostEngine, host_matrices)

_LIB = None


def library():
    """The product's native libraries (loaded once)."""
    global _LIB
    if _LIB is None:
        _LIB = Library()
    return _LIB
