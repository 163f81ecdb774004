"""ctypes binding of oracle/liboracle.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs
may import this module (see oracle/klt_oracle.h).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
REF_PATH = os.path.join(HERE, "_ref", "libklt_ref.so")
REF_QSORT_PATH = os.path.join(HERE, "_ref", "libklt_ref_qsort.so")

SORT_STABLE, SORT_QUICK = 0, 1


class Params(C.Structure):
    _fields_ = [
        ("mindist", C.c_int), ("window_width", C.c_int), ("window_height", C.c_int),
        ("smoothBeforeSelecting", C.c_int), ("min_eigenvalue", C.c_int),
        ("min_determinant", C.c_float), ("min_displacement", C.c_float),
        ("max_iterations", C.c_int), ("max_residue", C.c_float),
        ("grad_sigma", C.c_float), ("smooth_sigma_fact", C.c_float),
        ("pyramid_sigma_fact", C.c_float), ("step_factor", C.c_float),
        ("nSkippedPixels", C.c_int), ("borderx", C.c_int), ("bordery", C.c_int),
        ("nPyramidLevels", C.c_int), ("subsampling", C.c_int),
        ("lighting_insensitive", C.c_int),
    ]


class AffineParams(C.Structure):
    """klto_affine_params: the affine* members of KLT_TrackingContextRec (defaults klt.c:33-39)."""
    _fields_ = [("check", C.c_int), ("window_width", C.c_int), ("window_height", C.c_int),
                ("max_iterations", C.c_int), ("max_residue", C.c_float),
                ("min_displacement", C.c_float), ("max_displacement_differ", C.c_float)]


AFFINE_STATE = np.dtype([("has", "i4"), ("aff_x", "f4"), ("aff_y", "f4"), ("Axx", "f4"), ("Ayx", "f4"),
                         ("Axy", "f4"), ("Ayy", "f4")])


def affine_params(check=2, window=15, max_iterations=10, max_residue=10.0, min_displacement=0.02,
                  max_displacement_differ=1.5) -> AffineParams:
    return AffineParams(check, window, window, max_iterations, max_residue, min_displacement,
                        max_displacement_differ)


def affine_state(n, window=15):
    """fresh per-feature affine fields (KLTCreateFeatureList / selection: klt.c:160-176) + templates"""
    st = np.zeros(n, AFFINE_STATE)
    st["aff_x"] = st["aff_y"] = -1.0
    st["Axx"] = st["Ayy"] = 1.0
    return st, np.zeros((n, 3, (window + 2) * (window + 2)), np.float32)


def build(force: bool = False) -> None:
    """Compile liboracle.so (and oracle/_ref when /root/reference is present)."""
    if force or not os.path.exists(LIB_PATH) or \
            os.path.getmtime(LIB_PATH) < os.path.getmtime(os.path.join(HERE, "klt_oracle.c")):
        subprocess.check_call(["make", "-s", "-C", HERE, "liboracle.so"])
    if os.path.isdir("/root/reference/src/V3"):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")


class Oracle:
    def __init__(self, path: str = LIB_PATH):
        if not os.path.exists(path):
            build()
        L = self.lib = C.CDLL(path, mode=C.RTLD_LOCAL)
        P = C.POINTER(Params)
        L.klto_default_params.argtypes = [P]
        L.klto_change_pyramid.argtypes = [P, C.c_int]
        L.klto_update_border.argtypes = [P]
        L.klto_smooth_sigma.argtypes = [P]
        L.klto_smooth_sigma.restype = C.c_float
        L.klto_taps.argtypes = [C.c_float, _f32p, C.POINTER(C.c_int), _f32p, C.POINTER(C.c_int)]
        L.klto_taps.restype = C.c_int
        L.klto_to_float.argtypes = [_u8p, C.c_int, C.c_int, _f32p]
        L.klto_convolve_separate.argtypes = [_f32p, C.c_int, C.c_int, _f32p, C.c_int, _f32p, C.c_int, _f32p]
        L.klto_smooth.argtypes = [_f32p, C.c_int, C.c_int, C.c_float, _f32p]
        L.klto_gradients.argtypes = [_f32p, C.c_int, C.c_int, C.c_float, _f32p, _f32p]
        L.klto_pyr_down.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, C.c_float, _f32p]
        L.klto_mineig_points.argtypes = [_f32p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.c_int, C.c_int, C.c_int, _i32p]
        L.klto_mineig_points.restype = C.c_int
        L.klto_build_pyramids.argtypes = [_u8p, C.c_int, C.c_int, P]
        L.klto_build_pyramids.restype = C.c_void_p
        L.klto_free_pyramids.argtypes = [C.c_void_p]
        L.klto_pyr_levels.argtypes = [C.c_void_p]
        L.klto_pyr_levels.restype = C.c_int
        L.klto_pyr_dims.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.klto_pyr_data.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.klto_pyr_data.restype = C.POINTER(C.c_float)
        L.klto_track.argtypes = [C.c_void_p, C.c_void_p, P, C.c_int, _f32p, _f32p, _i32p]
        L.klto_track_level.argtypes = [C.c_float, C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                       _f32p, _f32p, _f32p, _f32p, _f32p, _f32p,
                                       C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                       C.c_int, C.c_float, C.c_float, C.c_float]
        L.klto_track_level.restype = C.c_int
        L.klto_track_affine.argtypes = [C.c_void_p, C.c_void_p, P, C.POINTER(AffineParams), C.c_int,
                                        _f32p, _f32p, _i32p, C.c_void_p, _f32p]
        L.klto_select.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, P, C.c_int, C.c_int,
                                  C.c_int, _f32p, _f32p, _i32p]
        L.klto_sort_points.argtypes = [_i32p, C.c_int, C.c_int]
        L.klto_enforce_min_distance.argtypes = [_i32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                                C.c_int, C.c_int, _f32p, _f32p, _i32p]

    # -- parameters -----------------------------------------------------------
    def default_params(self, **over) -> Params:
        p = Params()
        self.lib.klto_default_params(C.byref(p))
        for k, v in over.items():
            setattr(p, k, v)
        return p

    def update_border(self, p: Params) -> None:
        self.lib.klto_update_border(C.byref(p))

    def change_pyramid(self, p: Params, search_range: int) -> None:
        self.lib.klto_change_pyramid(C.byref(p), search_range)

    def taps(self, sigma: float):
        g = np.zeros(71, np.float32)
        d = np.zeros(71, np.float32)
        wg, wd = C.c_int(0), C.c_int(0)
        rc = self.lib.klto_taps(sigma, g, C.byref(wg), d, C.byref(wd))
        if rc != 0:
            raise ValueError("sigma %g needs more than 71 taps" % sigma)
        return g[:wg.value].copy(), d[:wd.value].copy()

    # -- stages ---------------------------------------------------------------
    def to_float(self, img):
        img = np.ascontiguousarray(img, np.uint8)
        out = np.empty(img.shape, np.float32)
        self.lib.klto_to_float(img, img.shape[1], img.shape[0], out)
        return out

    def smooth(self, f, sigma):
        f = np.ascontiguousarray(f, np.float32)
        out = np.empty_like(f)
        self.lib.klto_smooth(f, f.shape[1], f.shape[0], sigma, out)
        return out

    def gradients(self, f, sigma):
        f = np.ascontiguousarray(f, np.float32)
        gx, gy = np.empty_like(f), np.empty_like(f)
        self.lib.klto_gradients(f, f.shape[1], f.shape[0], sigma, gx, gy)
        return gx, gy

    def pyr_down(self, f, ss, sigma):
        f = np.ascontiguousarray(f, np.float32)
        out = np.empty((f.shape[0] // ss, f.shape[1] // ss), np.float32)
        self.lib.klto_pyr_down(f, f.shape[1], f.shape[0], ss, sigma, out)
        return out

    def mineig_points(self, gx, gy, ww, wh, bx, by, skip=0):
        pts = np.empty((gx.size, 3), np.int32)
        n = self.lib.klto_mineig_points(np.ascontiguousarray(gx), np.ascontiguousarray(gy),
                                        gx.shape[1], gx.shape[0], ww, wh, bx, by, skip, pts)
        return pts[:n].copy()

    def build_pyramids(self, img, p: Params):
        img = np.ascontiguousarray(img, np.uint8)
        return Pyramids(self, self.lib.klto_build_pyramids(img, img.shape[1], img.shape[0], C.byref(p)))

    def track(self, pyr1, pyr2, p: Params, x, y, val):
        x = np.array(x, np.float32); y = np.array(y, np.float32); val = np.array(val, np.int32)
        self.lib.klto_track(pyr1.handle, pyr2.handle, C.byref(p), len(x), x, y, val)
        return x, y, val

    def track_affine(self, pyr1, pyr2, p: Params, ap: AffineParams, x, y, val, state, tmpl):
        """KLTTrackFeatures with tc->affineConsistencyCheck >= 0; state / tmpl are updated in place."""
        x = np.array(x, np.float32); y = np.array(y, np.float32); val = np.array(val, np.int32)
        assert state.dtype == AFFINE_STATE and state.flags.c_contiguous and tmpl.flags.c_contiguous
        self.lib.klto_track_affine(pyr1.handle, pyr2.handle, C.byref(p), C.byref(ap), len(x), x, y, val,
                                   state.ctypes.data, tmpl.reshape(-1))
        return x, y, val

    def select(self, img, p: Params, n, sort_kind=SORT_STABLE, replace=False, last=None,
               x=None, y=None, val=None):
        img = np.ascontiguousarray(img, np.uint8)
        if x is None:
            x = np.zeros(n, np.float32); y = np.zeros(n, np.float32); val = np.zeros(n, np.int32)
        else:
            x = np.array(x, np.float32); y = np.array(y, np.float32); val = np.array(val, np.int32)
        self.lib.klto_select(img.ctypes.data, img.shape[1], img.shape[0],
                             last.handle if last is not None else None, C.byref(p),
                             sort_kind, 1 if replace else 0, n, x, y, val)
        return x, y, val


class Pyramids:
    def __init__(self, oracle: Oracle, handle):
        self.o, self.handle = oracle, handle

    @property
    def nlevels(self):
        return self.o.lib.klto_pyr_levels(self.handle)

    def level(self, which: int, l: int) -> np.ndarray:
        nc, nr = C.c_int(0), C.c_int(0)
        self.o.lib.klto_pyr_dims(self.handle, l, C.byref(nc), C.byref(nr))
        ptr = self.o.lib.klto_pyr_data(self.handle, which, l)
        return np.ctypeslib.as_array(ptr, shape=(nr.value, nc.value)).copy()

    def view(self, which: int, l: int) -> np.ndarray:
        """the level's data in place (writable, no copy): for sensitivity experiments in the tests"""
        nc, nr = C.c_int(0), C.c_int(0)
        self.o.lib.klto_pyr_dims(self.handle, l, C.byref(nc), C.byref(nr))
        return np.ctypeslib.as_array(self.o.lib.klto_pyr_data(self.handle, which, l), shape=(nr.value, nc.value))

    def __del__(self):
        try:
            self.o.lib.klto_free_pyramids(self.handle)
        except Exception:
            pass
