"""GPU tests of the affine consistency check (tc->affineConsistencyCheck = 0 / 1 / 2; reference
src/V1/trackFeatures.c:506-1224, :1438-1497; csrc/klt_affine.cuh) through the public KLTTrackFeatures.

The oracle's restatement is pinned bit-exact against the compiled reference in tests/test_oracle.py;
here the GPU path is compared with the oracle: in exact mode everything the reference leaves in the
feature list must be bit-identical (x, y, val, aff_x, aff_y, the map A, the saved template images).
"""
import ctypes as C

import numpy as np
import pytest

from tests.gpu_common import params_from_tc
from tests.test_oracle import _warped

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L(pkg):
    from importlib import import_module
    lib = import_module(pkg.__name__ + ".runtime").load()
    lib.require_gpu()
    lib.KLTSetVerbosity(0)
    return lib


class _FloatImage(C.Structure):
    _fields_ = [("ncols", C.c_int), ("nrows", C.c_int), ("data", C.POINTER(C.c_float))]


def _templates(fl, i):
    """the three saved template images of feature i as numpy arrays (None if absent)"""
    r = fl.contents.feature[i].contents
    if not r.aff_img:
        return None
    out = []
    for p in (r.aff_img, r.aff_img_gradx, r.aff_img_grady):
        im = C.cast(p, C.POINTER(_FloatImage)).contents
        out.append(np.ctypeslib.as_array(im.data, shape=(im.nrows * im.ncols,)).copy())
    return out


def _check_frame(capi, fl, x, y, v, st, tmpl, where):
    gx, gy, gv = capi.featurelist_to_arrays(fl)
    ga = capi.featurelist_affine(fl)
    assert np.array_equal(gv, v), "%s: status codes differ at %s" % (where, np.nonzero(gv != v)[0][:8])
    assert gx.tobytes() == x.tobytes() and gy.tobytes() == y.tobytes(), where
    assert np.array_equal(ga["has"], st["has"]), where
    live = st["has"] == 1
    for k in ("aff_x", "aff_y", "Axx", "Ayx", "Axy", "Ayy"):
        assert ga[k][live].tobytes() == st[k][live].tobytes(), (where, k)
    for i in np.nonzero(live)[0][:40]:
        t = _templates(fl, int(i))
        for w in range(3):
            assert t[w].tobytes() == tmpl[i, w].tobytes(), (where, int(i), w)


@pytest.mark.parametrize("check", [0, 1, 2])
@pytest.mark.parametrize("seq", ["provided", "warped"])
def test_affine_check_exact_mode_matches_oracle(L, capi, oracle, oracle_mod, provided, check, seq):
    imgs = provided[:7] if seq == "provided" else [_warped(provided[0], k) for k in range(7)]
    n = 120
    tc = L.KLTCreateTrackingContext()
    tc.contents.sequentialMode = 1
    tc.contents.affineConsistencyCheck = check
    L.KLTB200SetExact(tc, 1)
    fl = L.KLTCreateFeatureList(n)
    L.select(tc, imgs[0], fl)
    p = params_from_tc(oracle, tc)
    ap = oracle_mod.affine_params(check=check)
    x, y, v = oracle.select(imgs[0], p, n, sort_kind=oracle_mod.SORT_STABLE)
    st, tmpl = oracle_mod.affine_state(n)
    prev = oracle.build_pyramids(imgs[0], p)
    for i in range(1, len(imgs)):
        L.track(tc, imgs[i - 1], imgs[i], fl)
        cur = oracle.build_pyramids(imgs[i], p)
        x, y, v = oracle.track_affine(prev, cur, p, ap, x, y, v, st, tmpl)
        _check_frame(capi, fl, x, y, v, st, tmpl, "check %d %s frame %d" % (check, seq, i))
        prev = cur
    assert (st["has"] == 1).sum() > 20
    L.KLTFreeFeatureList(fl)
    L.KLTFreeTrackingContext(tc)


def test_affine_check_with_replacement_other_window_and_second_context(L, capi, oracle, oracle_mod, provided):
    """11x11 affine window, lost features replaced every frame (their affine members start over,
    selectGoodFeatures.c:514-541), and half way the list moves to a fresh tracking context: the
    templates then travel from the host images the library attached to the features."""
    n, check, win = 100, 2, 11
    imgs = provided[:8]

    def make_tc():
        tc = L.KLTCreateTrackingContext()
        tc.contents.sequentialMode = 1
        tc.contents.affineConsistencyCheck = check
        tc.contents.affine_window_width = win
        tc.contents.affine_window_height = win
        L.KLTB200SetExact(tc, 1)
        return tc

    tc = make_tc()
    fl = L.KLTCreateFeatureList(n)
    L.select(tc, imgs[0], fl)
    p = params_from_tc(oracle, tc)
    ap = oracle_mod.affine_params(check=check, window=win)
    x, y, v = oracle.select(imgs[0], p, n, sort_kind=oracle_mod.SORT_STABLE)
    st, tmpl = oracle_mod.affine_state(n, window=win)
    prev = oracle.build_pyramids(imgs[0], p)
    for i in range(1, len(imgs)):
        if i == 4:                                  # a new context: no device copies of the templates
            L.KLTFreeTrackingContext(tc)
            tc = make_tc()
        L.track(tc, imgs[i - 1], imgs[i], fl)
        cur = oracle.build_pyramids(imgs[i], p)
        x, y, v = oracle.track_affine(prev, cur, p, ap, x, y, v, st, tmpl)
        _check_frame(capi, fl, x, y, v, st, tmpl, "frame %d" % i)
        L.replace(tc, imgs[i], fl)
        lost = v < 0
        x, y, v = oracle.select(imgs[i], p, n, sort_kind=oracle_mod.SORT_STABLE, replace=True, last=cur,
                                x=x, y=y, val=v)
        st["has"][lost] = 0
        st["aff_x"][lost] = st["aff_y"][lost] = -1.0
        st["Axx"][lost] = st["Ayy"][lost] = 1.0
        st["Ayx"][lost] = st["Axy"][lost] = 0.0
        gx, gy, gv = capi.featurelist_to_arrays(fl)
        assert np.array_equal(gv, v) and gx.tobytes() == x.tobytes() and gy.tobytes() == y.tobytes()
        prev = cur
    L.KLTFreeFeatureList(fl)
    L.KLTFreeTrackingContext(tc)


def test_affine_check_fma_mode_agrees_with_exact_mode(L, capi, provided):
    """default (fma) arithmetic for the pyramids and the translation tracker, the same exact affine
    kernel behind it: status codes agree with the exact run on >= 97 % of the features after six
    free-running frames, coordinates of the common survivors within 0.01 px."""
    out = []
    for exact in (1, 0):
        tc = L.KLTCreateTrackingContext()
        tc.contents.sequentialMode = 1
        tc.contents.affineConsistencyCheck = 2
        L.KLTB200SetExact(tc, exact)
        fl = L.KLTCreateFeatureList(150)
        L.select(tc, provided[0], fl)
        for i in range(1, 7):
            L.track(tc, provided[i - 1], provided[i], fl)
        out.append(capi.featurelist_to_arrays(fl))
        L.KLTFreeFeatureList(fl)
        L.KLTFreeTrackingContext(tc)
    (ex, ey, ev), (fx, fy, fv) = out
    assert (ev == fv).mean() >= 0.97, "status agreement %.3f" % (ev == fv).mean()
    both = (ev >= 0) & (fv >= 0)
    assert both.sum() > 50
    assert max(np.abs(ex[both] - fx[both]).max(), np.abs(ey[both] - fy[both]).max()) <= 0.01


@pytest.mark.parametrize("replace", [0, 1])
@pytest.mark.parametrize("check,exact", [(2, 1), (1, 1), (0, 1), (2, 0)])
def test_affine_check_inside_the_sequence_call(L, capi, provided, check, exact, replace):
    """KLTTrackFeaturesSequence with tc->affineConsistencyCheck >= 0: the per-feature state (template
    held or not, aff_x / aff_y, the map A) and the templates stay on the device for the whole call and
    come back at its end.  Everything the call leaves behind -- the feature table, the final list, its
    affine members and template images -- is what the per-call loop leaves (reference driver loop
    src/V3/example3.c:54-76 around trackFeatures.c:1438-1497), bit for bit, with and without
    KLTReplaceLostFeatures after every frame (a refilled slot starts a new feature)."""
    n, imgs = 150, provided[:9]

    def run(sequence):
        tc = L.KLTCreateTrackingContext()
        tc.contents.sequentialMode = 1
        tc.contents.affineConsistencyCheck = check
        L.KLTB200SetExact(tc, exact)
        fl = L.KLTCreateFeatureList(n)
        ft = L.KLTCreateFeatureTable(len(imgs), n)
        L.select(tc, imgs[0], fl)
        L.KLTStoreFeatureList(fl, ft, 0)
        if sequence:
            L.track_sequence(tc, imgs[:5], fl, ft, 0, replace)       # two calls: the state travels through
            L.track_sequence(tc, imgs[4:], fl, ft, 4, replace)       # the list's host images in between
        else:
            for k in range(1, len(imgs)):
                L.track(tc, imgs[k - 1], imgs[k], fl)
                if replace:
                    L.replace(tc, imgs[k], fl)
                L.KLTStoreFeatureList(fl, ft, k)
        out = (capi.featuretable_to_array(ft), capi.featurelist_to_arrays(fl), capi.featurelist_affine(fl),
               [_templates(fl, i) for i in range(n)])
        L.KLTFreeFeatureTable(ft)
        L.KLTFreeFeatureList(fl)
        L.KLTFreeTrackingContext(tc)
        return out

    (ta, la, aa, ma), (tb, lb, ab, mb) = run(True), run(False)
    assert ta.tobytes() == tb.tobytes(), "feature tables differ in %d cells" % int((ta != tb).sum())
    for u, v in zip(la, lb):
        assert u.tobytes() == v.tobytes()
    assert np.array_equal(aa["has"], ab["has"])
    for k in ("aff_x", "aff_y", "Axx", "Ayx", "Axy", "Ayy"):
        assert aa[k].tobytes() == ab[k].tobytes(), k
    for i in range(n):
        assert (ma[i] is None) == (mb[i] is None), i
        if ma[i] is not None:
            for w in range(3):
                assert ma[i][w].tobytes() == mb[i][w].tobytes(), (i, w)
    assert (aa["has"] == 1).sum() > 20


@pytest.mark.parametrize("sequence", [False, True])
def test_affine_check_0_with_lighting_insensitive(L, capi, oracle, oracle_mod, provided, sequence):
    """affineConsistencyCheck = 0 together with lighting_insensitive (reference trackFeatures.c:1024-1028):
    the translation refinement against the template runs on gain / bias normalised windows.  Exact mode,
    frames with a growing brightness gain and offset: bit-identical to the oracle (pinned against the
    compiled reference in tests/test_oracle.py), through the per-call API and through the sequence call."""
    from tests.test_oracle import _ramped
    imgs = _ramped(provided, 7)
    n = 120
    tc = L.KLTCreateTrackingContext()
    tc.contents.sequentialMode = 1
    tc.contents.affineConsistencyCheck = 0
    tc.contents.lighting_insensitive = 1
    L.KLTB200SetExact(tc, 1)
    fl = L.KLTCreateFeatureList(n)
    L.select(tc, imgs[0], fl)
    p = params_from_tc(oracle, tc)
    assert p.lighting_insensitive == 1
    ap = oracle_mod.affine_params(check=0)
    x, y, v = oracle.select(imgs[0], p, n, sort_kind=oracle_mod.SORT_STABLE)
    st, tmpl = oracle_mod.affine_state(n)
    prev = oracle.build_pyramids(imgs[0], p)
    if sequence:
        L.track_sequence(tc, imgs, fl, None, 0, False)
    for i in range(1, len(imgs)):
        if not sequence:
            L.track(tc, imgs[i - 1], imgs[i], fl)
        cur = oracle.build_pyramids(imgs[i], p)
        x, y, v = oracle.track_affine(prev, cur, p, ap, x, y, v, st, tmpl)
        if not sequence:
            _check_frame(capi, fl, x, y, v, st, tmpl, "frame %d" % i)
        prev = cur
    _check_frame(capi, fl, x, y, v, st, tmpl, "end")
    assert (st["has"] == 1).sum() > 20
    L.KLTFreeFeatureList(fl)
    L.KLTFreeTrackingContext(tc)
