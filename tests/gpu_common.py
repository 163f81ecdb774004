"""Helpers shared by the GPU parity tests (all calls go through the C-ABI)."""
import ctypes as C

import numpy as np

# north_star tolerances
REL_TOL_IMAGES = 1e-4     # smoothed images, gradients, eigenvalue map (relative)
PX_TOL = 0.01             # tracked coordinates
STATUS_AGREE = 0.995      # fraction of features whose status code must agree


def rel_err(a, b):
    """|a-b| / max(|b|, 1)  (avoids blow-up near 0)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), 1.0)


def make_tc(L, oracle_params=None, **fields):
    tc = L.KLTCreateTrackingContext()
    for k, v in fields.items():
        setattr(tc.contents, k, v)
    return tc


def params_from_tc(oracle, tc):
    """oracle Params mirroring a KLT_TrackingContext field by field."""
    p = oracle.default_params()
    t = tc.contents
    for f in ("mindist", "window_width", "window_height", "smoothBeforeSelecting", "min_eigenvalue",
              "min_determinant", "min_displacement", "max_iterations", "max_residue", "grad_sigma",
              "smooth_sigma_fact", "pyramid_sigma_fact", "step_factor", "nSkippedPixels",
              "borderx", "bordery", "nPyramidLevels", "subsampling"):
        setattr(p, f, getattr(t, f))
    return p


def device_pyramids(L, dev, slot, nlevels):
    return [[L.dev_level(dev, slot, which, l) for l in range(nlevels)] for which in range(3)]


def compare_status(gx, gy, gv, ox, oy, ov):
    """-> (status agreement fraction, max coordinate error among features both track)"""
    agree = (gv == ov)
    both = agree & (ov >= 0)
    err = 0.0
    if both.any():
        err = float(max(np.abs(gx[both] - ox[both]).max(), np.abs(gy[both] - oy[both]).max()))
    return float(agree.mean()), err
