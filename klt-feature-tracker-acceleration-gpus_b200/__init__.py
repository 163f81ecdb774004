"""B200-native KLT tracker hot path behind the reference's C API.

The product is `lib/libklt_b200.so`: plain-C host code (the KLT public API of
the reference, include/klt.h) on top of a thin C-ABI (include/klt_cuda.h) into
hand-written sm_100a CUDA kernels (csrc/klt_dev.cu).  This Python package is
only a ctypes mirror of that C API for tests and benches; there is no Python or
CPU implementation of the path, and importing `.runtime` fails loudly when the
native library has not been built.
"""
from . import capi  # noqa: F401

__all__ = ["capi"]
