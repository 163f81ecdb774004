"""Out-of-plane stores, checked without compute-sanitizer (closed on the GPU pool).

Guard mode (klt_dev_set_guard, include/klt_cuda.h) re-allocates the pyramid arena with a 4 KB canary
band in front of every plane and behind the last one.  Every image kernel -- the three level-0
kernels, the fused level kernels in each tile shape, the tiled and the generic fallbacks -- then builds
awkward shapes (widths and heights that are no multiple of a tile, a strip, a segment or the pitch;
frames smaller than one tile) in both arithmetic modes, the results are compared with the oracle as in
tests/test_gpu_stages.py, and no canary word may have changed.  Selection, tracking, replacement and the
sequence call run on a guarded arena as well."""
import ctypes as C

import numpy as np
import pytest

from tests.conftest import synth_image
from tests.test_gpu_stages import _check_build

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L(pkg):
    from importlib import import_module
    lib = import_module(pkg.__name__ + ".runtime").load()
    lib.require_gpu()
    lib.KLTSetVerbosity(0)
    return lib


def _guards_intact(L, dev, where):
    bad, band = C.c_longlong(-1), C.c_int(-1)
    assert L.klt_dev_check_guards(dev, C.byref(bad), C.byref(band)) == 0
    assert bad.value == 0, "%s: %d canary words damaged, first in band %d" % (where, bad.value, band.value)


SHAPES = [(243, 321), (37, 1000), (600, 33), (130, 257), (65, 129), (200, 16), (40, 44), (481, 641), (95, 1217)]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("l0", [0, 1, 2])
def test_level0_kernels_stay_inside_their_planes(L, oracle, shape, l0):
    h, w = shape
    img = synth_image(w, h, seed=3 * h + w)
    before = L.klt_dev_l0_kernel()
    L.klt_dev_set_l0_kernel(l0)
    try:
        for exact in (1, 0):
            tc = L.KLTCreateTrackingContext()
            dev = L.KLTB200Device(tc)
            L.klt_dev_set_guard(dev, 1)
            _check_build(L, oracle, img, tc, exact, 0)
            _guards_intact(L, dev, "l0 kernel %d, %dx%d, exact %d" % (l0, w, h, exact))
            L.KLTFreeTrackingContext(tc)
    finally:
        L.klt_dev_set_l0_kernel(before)


@pytest.mark.parametrize("shape", [(243, 321), (700, 1100), (1081, 1923), (130, 257)])
@pytest.mark.parametrize("levels,ss", [(4, 2), (3, 4), (2, 4)])
@pytest.mark.parametrize("generic", [0, 1, 2])
def test_level_kernels_stay_inside_their_planes(L, oracle, shape, levels, ss, generic):
    h, w = shape
    if (w // ss ** (levels - 1)) < 8 or (h // ss ** (levels - 1)) < 8:
        pytest.skip("coarsest level too small for the default window")
    img = synth_image(w, h, seed=h + 7 * w)
    tc = L.KLTCreateTrackingContext()
    tc.contents.nPyramidLevels, tc.contents.subsampling = levels, ss
    L.KLTUpdateTCBorder(tc)
    dev = L.KLTB200Device(tc)
    L.klt_dev_set_guard(dev, 1)
    _check_build(L, oracle, img, tc, 0, generic)
    _guards_intact(L, dev, "%dx%d, %d levels, ss %d, path %d" % (w, h, levels, ss, generic))
    L.KLTFreeTrackingContext(tc)


def test_select_track_replace_sequence_on_a_guarded_arena(L, capi):
    frames = [synth_image(643, 487, 5, shift=(1.9 * k, -1.3 * k)) for k in range(8)]
    tc = L.KLTCreateTrackingContext()
    tc.contents.sequentialMode = 1
    tc.contents.nPyramidLevels, tc.contents.subsampling = 3, 2
    L.KLTUpdateTCBorder(tc)
    dev = L.KLTB200Device(tc)
    L.klt_dev_set_guard(dev, 1)
    fl = L.KLTCreateFeatureList(300)
    ft = L.KLTCreateFeatureTable(len(frames), 300)
    L.select(tc, frames[0], fl)
    for k in (1, 2, 3):
        L.track(tc, frames[k - 1], frames[k], fl)
        L.replace(tc, frames[k], fl)
    _guards_intact(L, dev, "per-call loop")
    L.track_sequence(tc, frames[3:], fl, ft, 3, True)
    _guards_intact(L, dev, "sequence call")
    assert (capi.featurelist_to_arrays(fl)[2] >= 0).sum() > 200
    L.KLTFreeFeatureTable(ft)
    L.KLTFreeFeatureList(fl)
    L.KLTFreeTrackingContext(tc)


@pytest.mark.parametrize("shape,n,mindist,window", [((243, 321), 500, 10, 7), ((97, 1217), 2000, 3, 5),
                                                     ((480, 640), 4000, 1, 7), ((1081, 1923), 1024, 25, 9),
                                                     ((130, 257), 50, 40, 11)])
def test_selection_buffers_stay_inside_their_allocations(L, capi, shape, n, mindist, window):
    """the eigenvalue map, the candidate lists of the radix sort, the rank list of the uncovered filter,
    the minimum-distance map (stamped as 32-bit words), the open-slot list and the feature arrays, each
    between two canary bands: selection, tracking and replacement with more features than the frame
    holds, tiny and huge minimum distances, other windows"""
    h, w = shape
    frames = [synth_image(w, h, 9, shift=(1.1 * k, 0.7 * k)) for k in range(3)]
    tc = L.KLTCreateTrackingContext()
    tc.contents.sequentialMode = 1
    tc.contents.mindist = mindist
    tc.contents.window_width = tc.contents.window_height = window
    L.KLTUpdateTCBorder(tc)
    dev = L.KLTB200Device(tc)
    L.klt_dev_set_guard(dev, 1)
    fl = L.KLTCreateFeatureList(n)
    L.select(tc, frames[0], fl)
    _guards_intact(L, dev, "select")
    for k in (1, 2):
        L.track(tc, frames[k - 1], frames[k], fl)
        x, y, v = capi.featurelist_to_arrays(fl)
        v = v.copy(); x = x.copy(); y = y.copy()
        v[::7] = -1; x[::7] = -1.0; y[::7] = -1.0               # open every seventh slot
        capi.arrays_to_featurelist(fl, x, y, v)
        L.replace(tc, frames[k], fl)
        _guards_intact(L, dev, "track + replace %d" % k)
    L.KLTFreeFeatureList(fl)
    L.KLTFreeTrackingContext(tc)


def test_a_store_outside_a_plane_is_seen(L, oracle):
    """the check itself: one float written right behind level 0's grady plane is reported"""
    img = synth_image(320, 240, seed=1)
    tc = L.KLTCreateTrackingContext()
    dev = L.KLTB200Device(tc)
    L.klt_dev_set_guard(dev, 1)
    _check_build(L, oracle, img, tc, 0, 0)
    _guards_intact(L, dev, "fresh")
    base = L.klt_dev_plane_address(dev, 0, 2, 0)           # slot 0, grady, level 0
    assert base
    pitch = (320 + 31) // 32 * 32
    one = C.c_float(0.0)
    assert L.klt_dev_poke(dev, C.c_void_p(base + pitch * 240 * 4), C.byref(one), 4) == 0
    bad, band = C.c_longlong(0), C.c_int(-1)
    assert L.klt_dev_check_guards(dev, C.byref(bad), C.byref(band)) == 0
    assert bad.value == 1 and band.value == 3              # the band in front of slot 0 / level 1 / image
    L.KLTFreeTrackingContext(tc)
