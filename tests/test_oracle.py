"""Pin the oracle (oracle/klt_oracle.c) before trusting it.

 * against the reference's golden run src/V1/feat/features2.ft (byte for byte)
 * against the unmodified reference CPU sources compiled in place (oracle/_ref),
   stage by stage and bit for bit.
CPU only.
"""
import ctypes as C

import numpy as np
import pytest

from tests.conftest import synth_image


def _v1_flow_oracle(oracle, oracle_mod, imgs, n=150, sort_kind=None):
    """reference src/V1/example3.c:35-65 restated on the oracle API."""
    sort_kind = oracle_mod.SORT_QUICK if sort_kind is None else sort_kind
    p = oracle.default_params()
    nframes = len(imgs)
    table = np.zeros((n, nframes), dtype=[("x", "f4"), ("y", "f4"), ("val", "i4")])
    x, y, v = oracle.select(imgs[0], p, n, sort_kind=sort_kind)
    table["x"][:, 0], table["y"][:, 0], table["val"][:, 0] = x, y, v
    prev = oracle.build_pyramids(imgs[0], p)
    for i in range(1, nframes):
        cur = oracle.build_pyramids(imgs[i], p)
        x, y, v = oracle.track(prev, cur, p, x, y, v)
        table["x"][:, i - 1], table["y"][:, i - 1], table["val"][:, i - 1] = x, y, v
        prev = cur
    return table


def test_oracle_reproduces_golden_feature_table(oracle, oracle_mod, provided, golden_ft):
    raw, gold = golden_ft
    tab = _v1_flow_oracle(oracle, oracle_mod, provided)
    # column 9 of the golden file was never written by the driver
    assert tab[:, :9].tobytes() == gold[:, :9].tobytes()
    # the whole file, header included, with the unwritten column taken as zeros
    tab[:, 9] = (0.0, 0.0, 0)
    blob = b"KLTFT1" + np.array([10, 150], np.int32).tobytes() + tab.tobytes()
    assert blob == raw


def test_oracle_call_counts_match_gprof(oracle, provided):
    # src/V1/example3_analysis.txt: 52 224 _minEigenvalue calls = (320-48)*(240-48)
    p = oracle.default_params()
    f = oracle.smooth(oracle.to_float(provided[0]), 0.7)
    gx, gy = oracle.gradients(f, 1.0)
    pts = oracle.mineig_points(gx, gy, 7, 7, p.borderx, p.bordery)
    assert len(pts) == 52224


@pytest.mark.parametrize("sigma,wg,wd", [(0.7, 5, 5), (1.0, 7, 7), (1.8, 11, 13),
                                         (3.6, 21, 25), (7.2, 43, 51)])
def test_tap_widths(oracle, ref, sigma, wg, wd):
    g, d = oracle.taps(sigma)
    assert (len(g), len(d)) == (wg, wd)
    assert ref.kernel_widths(sigma) == (wg, wd)
    assert abs(g.sum() - 1.0) < 1e-6
    assert d[len(d) // 2] == 0.0


def test_params_match_reference(oracle, ref):
    tc = ref.make_tc()
    p = oracle.default_params()
    for f in ("mindist", "window_width", "window_height", "min_eigenvalue", "max_iterations",
              "nSkippedPixels", "borderx", "bordery", "nPyramidLevels", "subsampling"):
        assert getattr(p, f) == getattr(tc.contents, f), f
    for sr in (1, 3, 5, 10, 15, 17, 20, 31, 32, 60, 100, 400):
        ref.api.KLTChangeTCPyramid(tc, sr)
        ref.api.KLTUpdateTCBorder(tc)
        oracle.change_pyramid(p, sr)
        oracle.update_border(p)
        assert (p.nPyramidLevels, p.subsampling, p.borderx, p.bordery) == \
            (tc.contents.nPyramidLevels, tc.contents.subsampling, tc.contents.borderx,
             tc.contents.bordery), sr
    # config 4 of BASELINE.json: L=4, ss=2 set directly, then the border update
    tc.contents.nPyramidLevels, tc.contents.subsampling = 4, 2
    ref.api.KLTUpdateTCBorder(tc)
    p.nPyramidLevels, p.subsampling = 4, 2
    oracle.update_border(p)
    assert p.borderx == tc.contents.borderx == 64
    ref.api.KLTFreeTrackingContext(tc)


@pytest.mark.parametrize("shape", [(243, 321), (64, 80), (31, 47)])
@pytest.mark.parametrize("sigma", [0.7, 1.0, 1.8, 3.6])
def test_stages_bit_exact_vs_reference(oracle, ref, shape, sigma):
    h, w = shape
    f = synth_image(w, h, seed=w * 7 + h).astype(np.float32)
    assert np.array_equal(oracle.smooth(f, sigma), ref.smooth(f, sigma))
    ogx, ogy = oracle.gradients(f, sigma)
    rgx, rgy = ref.gradients(f, sigma)
    assert np.array_equal(ogx, rgx) and np.array_equal(ogy, rgy)


@pytest.mark.parametrize("ss,levels", [(2, 4), (4, 2), (4, 3), (8, 2)])
def test_pyramid_bit_exact_vs_reference(oracle, ref, ss, levels):
    f = synth_image(352, 288, seed=ss * 10 + levels).astype(np.float32)
    want = ref.pyramid(f, ss, levels, 0.9)
    cur = f
    for l in range(1, levels):
        cur = oracle.pyr_down(cur, ss, np.float32(ss * np.float32(0.9)))
        assert cur.shape == want[l].shape
        assert np.array_equal(cur, want[l]), l


def _ref_flow(ref, imgs, n, **tc_fields):
    tc = ref.make_tc(sequentialMode=1, **tc_fields)
    fl = ref.new_list(n)
    ref.api.select(tc, imgs[0], fl)
    out = [ref.get(fl)]
    for i in range(1, len(imgs)):
        ref.api.track(tc, imgs[i - 1], imgs[i], fl)
        out.append(ref.get(fl))
    ref.api.KLTFreeFeatureList(fl)
    ref.api.KLTFreeTrackingContext(tc)
    return out


def test_select_and_track_bit_exact_vs_reference(oracle, oracle_mod, ref, ref_qsort, provided):
    imgs = provided[:4]
    for lib, kind in ((ref, oracle_mod.SORT_QUICK), (ref_qsort, oracle_mod.SORT_STABLE)):
        want = _ref_flow(lib, imgs, 150)
        p = oracle.default_params()
        x, y, v = oracle.select(imgs[0], p, 150, sort_kind=kind)
        assert np.array_equal(x, want[0][0]) and np.array_equal(y, want[0][1])
        assert np.array_equal(v, want[0][2])
        prev = oracle.build_pyramids(imgs[0], p)
        for i in range(1, len(imgs)):
            cur = oracle.build_pyramids(imgs[i], p)
            x, y, v = oracle.track(prev, cur, p, x, y, v)
            assert x.tobytes() == want[i][0].tobytes()
            assert y.tobytes() == want[i][1].tobytes()
            assert np.array_equal(v, want[i][2])
            prev = cur


def test_four_level_config_and_replace_vs_reference(oracle, oracle_mod, ref_qsort):
    """config-4 shaped parameters (L=4, ss=2) on a small synthetic pair, plus the
    KLTReplaceLostFeatures path that reuses pyramid_last (sequentialMode)."""
    imgs = [synth_image(480, 360, seed=5, shift=(1.7 * t, -0.9 * t)) for t in range(3)]
    n = 300
    tc = ref_qsort.make_tc(sequentialMode=1)
    tc.contents.nPyramidLevels, tc.contents.subsampling = 4, 2
    ref_qsort.api.KLTUpdateTCBorder(tc)
    fl = ref_qsort.new_list(n)
    p = oracle.default_params()
    p.nPyramidLevels, p.subsampling = 4, 2
    oracle.update_border(p)
    assert p.borderx == tc.contents.borderx

    ref_qsort.api.select(tc, imgs[0], fl)
    x, y, v = oracle.select(imgs[0], p, n, sort_kind=oracle_mod.SORT_STABLE)
    rx, ry, rv = ref_qsort.get(fl)
    assert np.array_equal(x, rx) and np.array_equal(y, ry) and np.array_equal(v, rv)
    prev = oracle.build_pyramids(imgs[0], p)
    for i in (1, 2):
        ref_qsort.api.track(tc, imgs[i - 1], imgs[i], fl)
        cur = oracle.build_pyramids(imgs[i], p)
        x, y, v = oracle.track(prev, cur, p, x, y, v)
        rx, ry, rv = ref_qsort.get(fl)
        assert x.tobytes() == rx.tobytes() and y.tobytes() == ry.tobytes()
        assert np.array_equal(v, rv)
        # knock out some features so that the replacement has work to do
        v[::7] = -3; x[::7] = -1; y[::7] = -1
        ref_qsort.put(fl, x, y, v)
        ref_qsort.api.replace(tc, imgs[i], fl)
        x, y, v = oracle.select(imgs[i], p, n, sort_kind=oracle_mod.SORT_STABLE,
                                replace=True, last=cur, x=x, y=y, val=v)
        rx, ry, rv = ref_qsort.get(fl)
        assert np.array_equal(x, rx) and np.array_equal(y, ry) and np.array_equal(v, rv)
        prev = cur
    # pyramids kept by the reference in sequential mode == the oracle's
    rimg, rgx, rgy = ref_qsort.last_pyramids(tc)
    for l in range(4):
        assert np.array_equal(rimg[l], prev.level(0, l))
        assert np.array_equal(rgx[l], prev.level(1, l))
        assert np.array_equal(rgy[l], prev.level(2, l))
    ref_qsort.api.KLTFreeFeatureList(fl)
    ref_qsort.api.KLTFreeTrackingContext(tc)


def test_quicksort_permutation_matches_reference(oracle, ref):
    rng = np.random.default_rng(3)
    for n in (1, 2, 3, 17, 1000, 52224):
        pts = np.zeros((n, 3), np.int32)
        pts[:, 0] = np.arange(n)
        pts[:, 1] = rng.integers(0, 1000, n)
        pts[:, 2] = rng.integers(0, 40, n)          # many ties
        a = pts.copy(); b = pts.copy()
        oracle.lib.klto_sort_points(a, n, 1)
        ref.lib._quicksort(b.ctypes.data_as(C.c_void_p), n)
        assert np.array_equal(a, b)
        c = pts.copy()
        oracle.lib.klto_sort_points(c, n, 0)
        order = np.argsort(-pts[:, 2], kind="stable")
        assert np.array_equal(c, pts[order])


def test_lighting_insensitive_tracking_bit_exact_vs_reference(oracle, oracle_mod, ref_qsort, provided):
    """tc->lighting_insensitive = TRUE (trackFeatures.c:125-220): gain / bias normalised windows,
    including the reference's quirk that the gradient sum's gain is sqrt(mean(g1) / mean(g2)).
    A brightness ramp is added to the later frames so that the normalisation matters."""
    imgs = [provided[0]]
    for k in range(1, 4):
        f = provided[k].astype(np.float32) * (1.0 + 0.08 * k) + 6.0 * k
        imgs.append(np.clip(f, 0, 255).astype(np.uint8))
    want = _ref_flow(ref_qsort, imgs, 150, lighting_insensitive=1)
    plain = _ref_flow(ref_qsort, imgs, 150)
    assert any(not np.array_equal(a[0], b[0]) for a, b in zip(want[1:], plain[1:]))   # the flag matters
    p = oracle.default_params()
    p.lighting_insensitive = 1
    x, y, v = oracle.select(imgs[0], p, 150, sort_kind=oracle_mod.SORT_STABLE)
    prev = oracle.build_pyramids(imgs[0], p)
    for i in range(1, len(imgs)):
        cur = oracle.build_pyramids(imgs[i], p)
        x, y, v = oracle.track(prev, cur, p, x, y, v)
        assert x.tobytes() == want[i][0].tobytes()
        assert y.tobytes() == want[i][1].tobytes()
        assert np.array_equal(v, want[i][2])
        prev = cur


def _ramped(provided, n):
    """the provided frames with a brightness gain / offset growing from frame to frame"""
    imgs = [provided[0]]
    for k in range(1, n):
        f = provided[k].astype(np.float32) * (1.0 + 0.01 * k) + 0.5 * k
        imgs.append(np.clip(f, 0, 255).astype(np.uint8))
    return imgs


def test_affine_check_0_with_lighting_insensitive_bit_exact_vs_reference(oracle, oracle_mod, capi, ref_qsort, provided):
    """tc->affineConsistencyCheck = 0 together with tc->lighting_insensitive (trackFeatures.c:1024-1028):
    the translation refinement against the template runs on the gain / bias normalised windows, the
    final residue on the plain difference (:1196-1199)."""
    imgs = _ramped(provided, 7)
    n = 120
    R = ref_qsort
    tc = R.make_tc(sequentialMode=1, affineConsistencyCheck=0, lighting_insensitive=1)
    fl = R.new_list(n)
    R.api.select(tc, imgs[0], fl)
    p = oracle.default_params()
    p.lighting_insensitive = 1
    ap = oracle_mod.affine_params(check=0)
    x, y, v = oracle.select(imgs[0], p, n, sort_kind=oracle_mod.SORT_STABLE)
    st, tmpl = oracle_mod.affine_state(n)
    prev = oracle.build_pyramids(imgs[0], p)
    p_plain = oracle.default_params()
    differs = 0
    for i in range(1, len(imgs)):
        R.api.track(tc, imgs[i - 1], imgs[i], fl)
        cur = oracle.build_pyramids(imgs[i], p)
        st2, tmpl2 = st.copy(), tmpl.copy()
        plain = oracle.track_affine(prev, cur, p_plain, ap, x.copy(), y.copy(), v.copy(), st2, tmpl2)
        x, y, v = oracle.track_affine(prev, cur, p, ap, x, y, v, st, tmpl)
        differs += int((plain[2] != v).sum()) + int((plain[0] != x).sum())
        rx, ry, rv = R.get(fl)
        ra = capi.featurelist_affine(fl)
        assert np.array_equal(v, rv), "frame %d" % i
        assert x.tobytes() == rx.tobytes() and y.tobytes() == ry.tobytes()
        assert np.array_equal(st["has"], ra["has"])
        live = st["has"] == 1
        for k in ("aff_x", "aff_y", "Axx", "Ayx", "Axy", "Ayy"):
            assert st[k][live].tobytes() == ra[k][live].tobytes(), (i, k)
        prev = cur
    assert differs > 0                                  # the flag matters on these frames
    assert (st["has"] == 1).sum() > 20
    R.api.KLTFreeFeatureList(fl)
    R.api.KLTFreeTrackingContext(tc)


def _warped(img, k):
    """frame k of a slowly rotating / zooming / shifting copy of img (bilinear, numpy)"""
    h, w = img.shape
    a, sc = np.deg2rad(0.6 * k), 1.0 + 0.004 * k
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    cx, cy = w / 2.0, h / 2.0
    xs = (np.cos(a) * (xx - cx) - np.sin(a) * (yy - cy)) / sc + cx + 0.9 * k
    ys = (np.sin(a) * (xx - cx) + np.cos(a) * (yy - cy)) / sc + cy - 0.6 * k
    x0 = np.clip(np.floor(xs).astype(int), 0, w - 2); y0 = np.clip(np.floor(ys).astype(int), 0, h - 2)
    ax = np.clip(xs - x0, 0, 1); ay = np.clip(ys - y0, 0, 1)
    f = img.astype(np.float64)
    v = (f[y0, x0] * (1 - ax) * (1 - ay) + f[y0, x0 + 1] * ax * (1 - ay)
         + f[y0 + 1, x0] * (1 - ax) * ay + f[y0 + 1, x0 + 1] * ax * ay)
    return np.clip(np.rint(v), 0, 255).astype(np.uint8)


@pytest.mark.parametrize("check", [0, 1, 2])
@pytest.mark.parametrize("seq", ["provided", "warped"])
def test_affine_consistency_check_bit_exact_vs_reference(oracle, oracle_mod, capi, ref_qsort, provided, check, seq):
    """tc->affineConsistencyCheck = 0 / 1 / 2 (trackFeatures.c:506-1224, :1438-1497): template
    creation after the first successful track, translation / similarity / affine refinement against
    the template, status changes, and the persistent aff_* members of every feature."""
    imgs = provided[:7] if seq == "provided" else [_warped(provided[0], k) for k in range(7)]
    n = 120
    R = ref_qsort
    tc = R.make_tc(sequentialMode=1, affineConsistencyCheck=check)
    fl = R.new_list(n)
    R.api.select(tc, imgs[0], fl)
    p = oracle.default_params()
    ap = oracle_mod.affine_params(check=check)
    x, y, v = oracle.select(imgs[0], p, n, sort_kind=oracle_mod.SORT_STABLE)
    st, tmpl = oracle_mod.affine_state(n)
    prev = oracle.build_pyramids(imgs[0], p)
    changed = 0
    for i in range(1, len(imgs)):
        R.api.track(tc, imgs[i - 1], imgs[i], fl)
        cur = oracle.build_pyramids(imgs[i], p)
        plain = oracle.track(prev, cur, p, x, y, v)
        x, y, v = oracle.track_affine(prev, cur, p, ap, x, y, v, st, tmpl)
        changed += int((plain[2] != v).sum())
        rx, ry, rv = R.get(fl)
        ra = capi.featurelist_affine(fl)
        assert np.array_equal(v, rv), "frame %d" % i
        assert x.tobytes() == rx.tobytes() and y.tobytes() == ry.tobytes()
        assert np.array_equal(st["has"], ra["has"])
        live = st["has"] == 1
        for k in ("aff_x", "aff_y", "Axx", "Ayx", "Axy", "Ayy"):
            assert st[k][live].tobytes() == ra[k][live].tobytes(), (i, k)
        prev = cur
    print("affine check %d on %s: %d status changes" % (check, seq, changed))
    R.api.KLTFreeFeatureList(fl)
    R.api.KLTFreeTrackingContext(tc)
