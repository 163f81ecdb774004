"""GPU tests of the batched driver API, KLTTrackFeaturesSequence (include/klt_b200.h).

The call must equal, frame for frame and bit for bit, the reference's driver loop
(src/V3/example3.c:54-76: KLTTrackFeatures [+ KLTReplaceLostFeatures] + KLTStoreFeatureList) run
through this library's per-call API, which the other GPU tests pin against the oracle; the exact-mode
cases are also checked against the oracle directly.
"""
import ctypes as C

import numpy as np
import pytest

from tests.conftest import synth_image
from tests.gpu_common import params_from_tc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L(pkg):
    from importlib import import_module
    lib = import_module(pkg.__name__ + ".runtime").load()
    lib.require_gpu()
    lib.KLTSetVerbosity(0)
    return lib


def _loop(L, capi, frames, n, exact, replace, tc_setup=None):
    """the per-call driver loop -> [nFeatures, nFrames] table, final list"""
    tc = L.KLTCreateTrackingContext()
    tc.contents.sequentialMode = 1
    if tc_setup:
        tc_setup(tc)
    L.KLTB200SetExact(tc, exact)
    fl = L.KLTCreateFeatureList(n)
    ft = L.KLTCreateFeatureTable(len(frames), n)
    L.select(tc, frames[0], fl)
    L.KLTStoreFeatureList(fl, ft, 0)
    for k in range(1, len(frames)):
        L.track(tc, frames[k - 1], frames[k], fl)
        if replace:
            L.replace(tc, frames[k], fl)
        L.KLTStoreFeatureList(fl, ft, k)
    tab = capi.featuretable_to_array(ft)
    fin = capi.featurelist_to_arrays(fl)
    L.KLTFreeFeatureTable(ft)
    L.KLTFreeFeatureList(fl)
    L.KLTFreeTrackingContext(tc)
    return tab, fin


def _sequence(L, capi, frames, n, exact, replace, tc_setup=None, split=None):
    tc = L.KLTCreateTrackingContext()
    tc.contents.sequentialMode = 1
    if tc_setup:
        tc_setup(tc)
    L.KLTB200SetExact(tc, exact)
    fl = L.KLTCreateFeatureList(n)
    ft = L.KLTCreateFeatureTable(len(frames), n)
    L.select(tc, frames[0], fl)
    L.KLTStoreFeatureList(fl, ft, 0)
    if split is None:
        L.track_sequence(tc, frames, fl, ft, 0, replace)
    else:                                   # two calls: the second continues from the held pyramids
        L.track_sequence(tc, frames[:split + 1], fl, ft, 0, replace)
        L.track_sequence(tc, frames[split:], fl, ft, split, replace)
    tab = capi.featuretable_to_array(ft)
    fin = capi.featurelist_to_arrays(fl)
    assert tc.contents.sequentialMode == 1 and tc.contents.pyramid_last
    L.KLTFreeFeatureTable(ft)
    L.KLTFreeFeatureList(fl)
    L.KLTFreeTrackingContext(tc)
    return tab, fin


def _same(a, b):
    ta, fa = a
    tb, fb = b
    assert ta.tobytes() == tb.tobytes(), "feature tables differ in %d cells" % int((ta != tb).sum())
    for u, v in zip(fa, fb):
        assert u.tobytes() == v.tobytes()


@pytest.mark.parametrize("exact", [1, 0])
@pytest.mark.parametrize("replace", [False, True])
def test_sequence_equals_driver_loop_config1(L, capi, provided, exact, replace):
    _same(_sequence(L, capi, provided, 150, exact, replace), _loop(L, capi, provided, 150, exact, replace))


def test_sequence_matches_oracle_exact(L, capi, oracle, oracle_mod, provided):
    n = 150
    tab, fin = _sequence(L, capi, provided, n, 1, True)
    tc = L.KLTCreateTrackingContext()
    p = params_from_tc(oracle, tc)
    L.KLTFreeTrackingContext(tc)
    ox, oy, ov = oracle.select(provided[0], p, n, sort_kind=oracle_mod.SORT_STABLE)
    prev = oracle.build_pyramids(provided[0], p)
    for k in range(1, len(provided)):
        cur = oracle.build_pyramids(provided[k], p)
        ox, oy, ov = oracle.track(prev, cur, p, ox, oy, ov)
        ox, oy, ov = oracle.select(provided[k], p, n, sort_kind=oracle_mod.SORT_STABLE,
                                   replace=True, last=cur, x=ox, y=oy, val=ov)
        assert np.array_equal(tab["val"][:, k], ov), "frame %d" % k
        assert tab["x"][:, k].tobytes() == ox.tobytes() and tab["y"][:, k].tobytes() == oy.tobytes()
        prev = cur
    assert np.array_equal(fin[2], ov)


def test_sequence_in_two_calls_and_four_levels(L, capi):
    """1080p, 4 pyramid levels (banded upload of every frame), continued by a second call."""
    frames = [synth_image(1920, 1080, 7, shift=(1.7 * k, -1.1 * k)) for k in range(6)]

    def setup(tc):
        tc.contents.nPyramidLevels = 4
        tc.contents.subsampling = 2
        L.KLTUpdateTCBorder(tc)

    ref = _loop(L, capi, frames, 1024, 0, False, setup)
    _same(_sequence(L, capi, frames, 1024, 0, False, setup), ref)
    _same(_sequence(L, capi, frames, 1024, 0, False, setup, split=3), ref)
    assert (ref[1][2] >= 0).sum() > 900


def test_sequence_without_table_and_single_frame(L, capi, provided):
    tc = L.KLTCreateTrackingContext()
    tc.contents.sequentialMode = 1
    fl = L.KLTCreateFeatureList(100)
    L.select(tc, provided[0], fl)
    before = capi.featurelist_to_arrays(fl)
    L.track_sequence(tc, provided[:1], fl)                     # nothing to track into
    after = capi.featurelist_to_arrays(fl)
    for u, v in zip(before, after):
        assert u.tobytes() == v.tobytes()
    L.track_sequence(tc, provided[:4], fl)                     # ft == NULL
    a = capi.featurelist_to_arrays(fl)
    L.KLTFreeFeatureList(fl)
    L.KLTFreeTrackingContext(tc)
    _, fin = _loop(L, capi, provided[:4], 100, 0, False)
    for u, v in zip(a, fin):
        assert u.tobytes() == v.tobytes()


@pytest.mark.parametrize("exact", [1, 0])
def test_sequence_with_replacement_in_overlap_mode(L, capi, provided, exact, monkeypatch):
    """KLT_B200_OVERLAP=1: the tracker and the feature copies run on a second stream, the selection
    of a replacement on the build stream.  The snapshot / final fetch that follow on the tracker
    stream must be ordered after the selection (an event join at the end of select_core): the
    tables must equal the in-order pipeline's bit for bit, every time."""
    ref = _sequence(L, capi, provided, 150, exact, True)
    monkeypatch.setenv("KLT_B200_OVERLAP", "1")
    for _ in range(3):
        _same(_sequence(L, capi, provided, 150, exact, True), ref)
    frames = [synth_image(1920, 1080, 11, shift=(1.3 * k, 0.9 * k)) for k in range(8)]
    monkeypatch.setenv("KLT_B200_OVERLAP", "0")
    big = _sequence(L, capi, frames, 2000, exact, True)
    monkeypatch.setenv("KLT_B200_OVERLAP", "1")
    _same(_sequence(L, capi, frames, 2000, exact, True), big)
