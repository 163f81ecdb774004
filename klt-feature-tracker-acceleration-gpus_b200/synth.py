"""Binding of lib/libklt_synth.so: deterministic synthetic sequences (BASELINE
configs 4/5).  Host-side data generation only -- not part of the hot path."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libklt_synth.so")
_lib = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s not built (make -C %s/csrc)" % (LIB_PATH, HERE))
        _lib = C.CDLL(LIB_PATH)
        _lib.klt_synth_frame.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint, C.c_float,
                                         C.c_float, C.c_float, C.c_float, C.c_float, C.c_int]
        _lib.klt_synth_frame.restype = None
    return _lib


def frame(ncols, nrows, seed=12345, t=0.0, velocity=(2.3, -1.4), rot_deg=0.0, scale=1.0,
          out=None, threads=None) -> np.ndarray:
    """One u8 frame of the sequence; `out` may be any writable C-contiguous uint8 [nrows, ncols]."""
    if out is None:
        out = np.empty((nrows, ncols), np.uint8)
    assert out.dtype == np.uint8 and out.shape == (nrows, ncols) and out.flags.c_contiguous
    if threads is None:
        threads = min(os.cpu_count() or 1, 32)
    _load().klt_synth_frame(out.ctypes.data, ncols, nrows, seed, float(t), float(velocity[0]),
                            float(velocity[1]), float(rot_deg), float(scale), int(threads))
    return out


def pingpong_index(step: int, nframes: int) -> int:
    """0,1,..,n-1,n-2,..,1,0,1,..: a finite set of frames traversed with bounded motion."""
    if nframes < 2:
        return 0
    period = 2 * (nframes - 1)
    p = step % period
    return p if p < nframes else period - p
