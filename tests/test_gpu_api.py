"""Parity of the public KLT API on the GPU (KLTSelectGoodFeatures,
KLTTrackFeatures, KLTReplaceLostFeatures) against the oracle.

exact mode: bit-identical feature lists, including the golden V1 run.
fast mode : teacher-forced per frame pair, <= 0.01 px and >= 99.5 % status
            agreement (north_star), disagreements listed.
"""
import ctypes as C

import numpy as np
import pytest

from tests.conftest import synth_image
from tests.gpu_common import PX_TOL, STATUS_AGREE, check_fma_step, compare_status, params_from_tc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L(pkg):
    from importlib import import_module
    lib = import_module(pkg.__name__ + ".runtime").load()
    lib.require_gpu()
    lib.KLTSetVerbosity(0)
    return lib


def _get(capi, fl):
    return capi.featurelist_to_arrays(fl)


def test_select_matches_oracle_stable_order(L, capi, oracle, oracle_mod, provided):
    tc = L.KLTCreateTrackingContext()
    fl = L.KLTCreateFeatureList(150)
    L.select(tc, provided[0], fl)
    x, y, v = _get(capi, fl)
    ox, oy, ov = oracle.select(provided[0], params_from_tc(oracle, tc), 150, sort_kind=oracle_mod.SORT_STABLE)
    assert np.array_equal(x, ox) and np.array_equal(y, oy) and np.array_equal(v, ov)
    assert (v > 0).all()
    L.KLTFreeFeatureList(fl)
    L.KLTFreeTrackingContext(tc)


@pytest.mark.parametrize("n,mindist,min_eig", [(1, 10, 1), (2000, 10, 1), (5000, 25, 1),
                                               (300, 0, 1), (300, 1, 500), (40000, 3, 1),
                                               (40000, 2, 1), (70000, 1, 1), (3000, 5, 1), (60, 60, 1)])
def test_select_edge_cases(L, capi, oracle, oracle_mod, provided, n, mindist, min_eig):
    """more features asked than exist (NOT_FOUND padding), mindist 0/1, thresholds"""
    tc = L.KLTCreateTrackingContext()
    tc.contents.mindist, tc.contents.min_eigenvalue = mindist, min_eig
    fl = L.KLTCreateFeatureList(n)
    L.select(tc, provided[3], fl)
    x, y, v = _get(capi, fl)
    ox, oy, ov = oracle.select(provided[3], params_from_tc(oracle, tc), n, sort_kind=oracle_mod.SORT_STABLE)
    assert np.array_equal(v, ov)
    assert np.array_equal(x, ox) and np.array_equal(y, oy)
    L.KLTFreeFeatureList(fl)
    L.KLTFreeTrackingContext(tc)


def test_golden_run_exact_mode(L, capi, oracle, oracle_mod, provided, golden_ft):
    """reference src/V1/example3.c flow, 150 features, 10 frames, sequentialMode.
    GPU exact mode == oracle (stable ranking) bit for bit; and == the golden file
    for every feature whose slot is not affected by the 2 tie swaps."""
    _, gold = golden_ft
    tc = L.KLTCreateTrackingContext()
    tc.contents.sequentialMode = 1
    L.KLTB200SetExact(tc, 1)
    fl = L.KLTCreateFeatureList(150)
    ft = L.KLTCreateFeatureTable(10, 150)
    L.select(tc, provided[0], fl)
    L.KLTStoreFeatureList(fl, ft, 0)
    p = params_from_tc(oracle, tc)
    ox, oy, ov = oracle.select(provided[0], p, 150, sort_kind=oracle_mod.SORT_STABLE)
    prev = oracle.build_pyramids(provided[0], p)
    for i in range(1, 10):
        L.track(tc, provided[i - 1], provided[i], fl)
        L.KLTStoreFeatureList(fl, ft, i - 1)
        cur = oracle.build_pyramids(provided[i], p)
        ox, oy, ov = oracle.track(prev, cur, p, ox, oy, ov)
        prev = cur
        x, y, v = _get(capi, fl)
        assert x.tobytes() == ox.tobytes() and y.tobytes() == oy.tobytes(), "frame %d" % i
        assert np.array_equal(v, ov)
    tab = capi.featuretable_to_array(ft)
    # ties: the golden file comes from the unstable _quicksort build; slots
    # 94<->95 and 144<->145 hold the same two features in swapped order.
    perm = np.arange(150)
    perm[[94, 95, 144, 145]] = [95, 94, 145, 144]
    assert tab[:, :9].tobytes() == gold[perm][:, :9].tobytes()
    L.KLTFreeFeatureTable(ft)
    L.KLTFreeFeatureList(fl)
    L.KLTFreeTrackingContext(tc)


def _teacher_forced(L, capi, oracle, oracle_mod, imgs, n, exact, tc_setup=None, no_fused=0):
    tc = L.KLTCreateTrackingContext()
    tc.contents.sequentialMode = 1
    if tc_setup:
        tc_setup(tc)
    L.KLTB200SetExact(tc, exact)
    L.klt_dev_disable_fused(L.KLTB200Device(tc), no_fused)
    p = params_from_tc(oracle, tc)
    fl = L.KLTCreateFeatureList(n)
    ox, oy, ov = oracle.select(imgs[0], p, n, sort_kind=oracle_mod.SORT_STABLE)
    prev = oracle.build_pyramids(imgs[0], p)
    report = []
    for i in range(1, len(imgs)):
        capi.arrays_to_featurelist(fl, ox, oy, ov)      # oracle state in (teacher forcing)
        L.track(tc, imgs[i - 1], imgs[i], fl)
        gx, gy, gv = _get(capi, fl)
        cur = oracle.build_pyramids(imgs[i], p)
        x0, y0, v0 = ox, oy, ov
        ox, oy, ov = oracle.track(prev, cur, p, ox, oy, ov)
        agree, err = compare_status(gx, gy, gv, ox, oy, ov)
        report.append((i, agree, err, int((ov >= 0).sum())))
        if exact:
            assert gx.tobytes() == ox.tobytes() and gy.tobytes() == oy.tobytes() and np.array_equal(gv, ov)
        else:
            check_fma_step(oracle, p, prev, cur, x0, y0, v0, gx, gy, gv, ox, oy, ov, "frame %d" % i)
        prev = cur
    L.KLTFreeFeatureList(fl)
    L.KLTFreeTrackingContext(tc)
    return report


@pytest.mark.parametrize("exact", [1, 0])
def test_track_teacher_forced_config1(L, capi, oracle, oracle_mod, provided, exact):
    rep = _teacher_forced(L, capi, oracle, oracle_mod, provided, 150, exact)
    assert rep[-1][3] > 50          # most features survive the 10 frames


@pytest.mark.parametrize("exact", [1, 0])
def test_track_teacher_forced_four_levels(L, capi, oracle, oracle_mod, exact):
    imgs = [synth_image(900, 700, seed=11, shift=(2.3 * t, -1.4 * t)) for t in range(4)]

    def setup(tc):
        tc.contents.nPyramidLevels, tc.contents.subsampling = 4, 2
        L.KLTUpdateTCBorder(tc)
    rep = _teacher_forced(L, capi, oracle, oracle_mod, imgs, 600, exact, setup)
    assert rep[-1][3] > 300


def test_synthetic_translation_is_recovered(L, capi):
    """T7: pure translation (2.3, -1.4) px per frame, fast mode."""
    imgs = [synth_image(800, 600, seed=21, shift=(2.3 * t, -1.4 * t)) for t in range(3)]
    tc = L.KLTCreateTrackingContext()
    tc.contents.sequentialMode = 1
    fl = L.KLTCreateFeatureList(400)
    L.select(tc, imgs[0], fl)
    x0, y0, v0 = _get(capi, fl)
    L.track(tc, imgs[0], imgs[1], fl)
    x1, y1, v1 = _get(capi, fl)
    ok = v1 >= 0
    assert ok.mean() > 0.9
    # the image content moves by -shift
    assert abs(np.median(x1[ok] - x0[ok]) + 2.3) < 0.1
    assert abs(np.median(y1[ok] - y0[ok]) - 1.4) < 0.1
    L.KLTFreeFeatureList(fl)
    L.KLTFreeTrackingContext(tc)


@pytest.mark.parametrize("exact", [1, 0])
def test_replace_lost_features_sequential(L, capi, oracle, oracle_mod, provided, exact):
    """KLTReplaceLostFeatures reuses the device-resident level 0 of the last tracked frame
    (reference selectGoodFeatures.c:342-348).  Its ranking keys are truncated integers, so in the
    default fma mode level 0 is rebuilt in exact arithmetic for it (klt_dev_exact_level0): from the
    oracle's tracked list the refilled list is bit-identical in BOTH modes."""
    n = 200
    tc = L.KLTCreateTrackingContext()
    tc.contents.sequentialMode = 1
    L.KLTB200SetExact(tc, exact)
    p = params_from_tc(oracle, tc)
    fl = L.KLTCreateFeatureList(n)
    L.select(tc, provided[0], fl)
    ox, oy, ov = oracle.select(provided[0], p, n, sort_kind=oracle_mod.SORT_STABLE)
    prev = oracle.build_pyramids(provided[0], p)
    for i in range(1, 5):
        L.track(tc, provided[i - 1], provided[i], fl)
        cur = oracle.build_pyramids(provided[i], p)
        ox, oy, ov = oracle.track(prev, cur, p, ox, oy, ov)
        if exact:
            x, y, v = _get(capi, fl)
            assert np.array_equal(v, ov) and x.tobytes() == ox.tobytes() and y.tobytes() == oy.tobytes()
        else:
            capi.arrays_to_featurelist(fl, ox, oy, ov)          # teacher forcing
        L.replace(tc, provided[i], fl)
        ox, oy, ov = oracle.select(provided[i], p, n, sort_kind=oracle_mod.SORT_STABLE,
                                   replace=True, last=cur, x=ox, y=oy, val=ov)
        x, y, v = _get(capi, fl)
        assert np.array_equal(v, ov), "frame %d" % i
        assert x.tobytes() == ox.tobytes() and y.tobytes() == oy.tobytes()
        prev = cur
    L.KLTFreeFeatureList(fl)
    L.KLTFreeTrackingContext(tc)


def test_lost_features_are_left_alone_and_nonsequential_mode(L, capi, oracle, oracle_mod, provided):
    tc = L.KLTCreateTrackingContext()          # sequentialMode FALSE: both pyramids every call
    L.KLTB200SetExact(tc, 1)
    p = params_from_tc(oracle, tc)
    n = 64
    fl = L.KLTCreateFeatureList(n)
    ox, oy, ov = oracle.select(provided[0], p, n, sort_kind=oracle_mod.SORT_STABLE)
    ov[::5] = -3; ox[::5] = -1; oy[::5] = -1
    capi.arrays_to_featurelist(fl, ox, oy, ov)
    L.track(tc, provided[0], provided[2], fl)
    assert tc.contents.pyramid_last is None
    x, y, v = _get(capi, fl)
    a = oracle.build_pyramids(provided[0], p)
    b = oracle.build_pyramids(provided[2], p)
    ox, oy, ov = oracle.track(a, b, p, ox, oy, ov)
    assert np.array_equal(v, ov) and x.tobytes() == ox.tobytes() and y.tobytes() == oy.tobytes()
    assert (v[::5] == -3).all()
    L.KLTFreeFeatureList(fl)
    L.KLTFreeTrackingContext(tc)


@pytest.mark.parametrize("window", [(3, 3), (5, 5), (9, 9), (11, 11), (15, 15), (7, 5), (5, 9), (17, 17)])
@pytest.mark.parametrize("exact", [1, 0])
def test_track_other_window_sizes(L, capi, oracle, oracle_mod, window, exact):
    """square windows up to 15 use the 8-lanes-per-feature kernel in fma mode, everything
    else (and the exact mode) the warp-per-feature kernel"""
    imgs = [synth_image(500, 400, seed=31, shift=(1.9 * t, 1.1 * t)) for t in range(3)]

    def setup(tc):
        tc.contents.window_width, tc.contents.window_height = window
        L.KLTChangeTCPyramid(tc, 12)
        L.KLTUpdateTCBorder(tc)
    rep = _teacher_forced(L, capi, oracle, oracle_mod, imgs, 200, exact, setup)
    assert rep[-1][3] > 60


def test_track_7x7_generic_fast_kernel(L, capi, oracle, oracle_mod, provided):
    """fma mode without the specialised kernels: tiled image kernels + track_fast_kernel<7,1>"""
    rep = _teacher_forced(L, capi, oracle, oracle_mod, provided[:5], 150, 0, no_fused=1)
    assert rep[-1][3] > 80


@pytest.mark.parametrize("kernel", ["track7w", "track_fast"])
def test_track_7x7_kernel_generations(L, capi, oracle, oracle_mod, kernel):
    """the two 7x7 fma trackers (one warp per feature: default; 8 lanes per feature) against the
    oracle, teacher-forced, 4 levels, on frames whose footprints take every alignment"""
    imgs = [synth_image(640, 480, seed=5, shift=(1.9 * t, -1.3 * t)) for t in range(5)]

    def setup(tc):
        tc.contents.nPyramidLevels, tc.contents.subsampling = 4, 2
        L.KLTUpdateTCBorder(tc)
        dev = L.KLTB200Device(tc)
        L.klt_dev_disable_track7w(dev, 0 if kernel == "track7w" else 1)

    rep = _teacher_forced(L, capi, oracle, oracle_mod, imgs, 400, 0, tc_setup=setup)
    assert rep[-1][3] > 250


def test_record_mode_equals_staging_mode(L, capi, provided, monkeypatch):
    """KLTCreateFeatureList pins its block when a device is present; KLTTrackFeatures then mirrors the
    records in one copy and the tracker writes x | y | val straight back into them.  A list in
    ordinary memory (KLT_B200_PINNED_LISTS=0) takes the pack / staging / unpack path.  Both must give
    the same features bit for bit, including the untouched lost ones."""
    import ctypes as C
    res = []
    for pinned in ("1", "0"):
        monkeypatch.setenv("KLT_B200_PINNED_LISTS", pinned)
        tc = L.KLTCreateTrackingContext()
        tc.contents.sequentialMode = 1
        fl = L.KLTCreateFeatureList(200)
        L.select(tc, provided[0], fl)
        x, y, v = capi.featurelist_to_arrays(fl)
        v[::7] = -3                                   # some features already lost: must stay as they are
        x[::7] = -1.0
        y[::7] = -1.0
        capi.arrays_to_featurelist(fl, x, y, v)
        for k in range(1, 4):
            L.track(tc, provided[k - 1], provided[k], fl)
        res.append(capi.featurelist_to_arrays(fl))
        L.KLTFreeFeatureList(fl)
        L.KLTFreeTrackingContext(tc)
    for a, b in zip(res[0], res[1]):
        assert np.array_equal(a, b)
    assert (res[0][2][::7] == -3).all() and (res[0][0][::7] == -1.0).all()


@pytest.mark.parametrize("window", [7, 5])
@pytest.mark.parametrize("exact", [1, 0])
def test_lighting_insensitive_tracking(L, capi, oracle, oracle_mod, provided, exact, window):
    """tc->lighting_insensitive = TRUE (reference trackFeatures.c:125-220, :433-437, :466-468) on the
    GPU: gain / bias normalised windows, with the reference's own gain formula for the gradient sum.
    Frames get brighter over time so that the normalisation matters; exact mode reproduces the
    oracle bit for bit, fma mode meets the north_star tolerance."""
    imgs = [provided[0]]
    for k in range(1, 5):
        f = provided[k].astype(np.float32) * (1.0 + 0.08 * k) + 6.0 * k
        imgs.append(np.clip(f, 0, 255).astype(np.uint8))

    def setup(tc):
        tc.contents.lighting_insensitive = 1
        tc.contents.window_width = tc.contents.window_height = window
    rep = _teacher_forced(L, capi, oracle, oracle_mod, imgs, 150, exact, setup)
    assert rep[-1][3] > 50
    # and it is a different computation from the plain tracker on these frames
    def plain(tc):
        tc.contents.window_width = tc.contents.window_height = window
    rep2 = _teacher_forced(L, capi, oracle, oracle_mod, imgs, 150, exact, plain)
    assert rep != rep2


# ---- edge cases ---------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(40, 44), (49, 49), (30, 200), (64, 52)])
def test_tiny_images_select_and_track(L, capi, oracle, oracle_mod, shape):
    """images barely larger than (or smaller than) twice the default border of 24: few or no
    candidates, every slot padded with NOT_FOUND exactly as the reference does, tracking a list with
    nothing alive is a no-op."""
    h, w = shape
    imgs = [synth_image(w, h, seed=5 + t, shift=(0.7 * t, 0.3 * t)) for t in range(2)]
    tc = L.KLTCreateTrackingContext()
    fl = L.KLTCreateFeatureList(20)
    L.select(tc, imgs[0], fl)
    x, y, v = _get(capi, fl)
    p = params_from_tc(oracle, tc)
    ox, oy, ov = oracle.select(imgs[0], p, 20, sort_kind=oracle_mod.SORT_STABLE)
    assert np.array_equal(v, ov) and np.array_equal(x, ox) and np.array_equal(y, oy)
    L.KLTB200SetExact(tc, 1)
    L.track(tc, imgs[0], imgs[1], fl)
    gx, gy, gv = _get(capi, fl)
    tx, ty, tv = oracle.track(oracle.build_pyramids(imgs[0], p), oracle.build_pyramids(imgs[1], p), p, ox, oy, ov)
    assert np.array_equal(gv, tv) and gx.tobytes() == tx.tobytes() and gy.tobytes() == ty.tobytes()
    L.KLTFreeFeatureList(fl)
    L.KLTFreeTrackingContext(tc)


def test_even_and_small_windows_are_repaired(L, capi, oracle, oracle_mod, provided):
    """window 6x4 -> 7x5 and 2x2 -> 3x3, silently, as trackFeatures.c:1258-1278 /
    selectGoodFeatures.c:313-333 do; results equal the oracle run with the repaired sizes."""
    for (ww, wh), (rw, rh) in (((6, 4), (7, 5)), ((2, 2), (3, 3))):
        tc = L.KLTCreateTrackingContext()
        tc.contents.window_width, tc.contents.window_height = ww, wh
        L.KLTB200SetExact(tc, 1)
        fl = L.KLTCreateFeatureList(100)
        L.select(tc, provided[0], fl)
        assert (tc.contents.window_width, tc.contents.window_height) == (rw, rh)
        p = params_from_tc(oracle, tc)
        ox, oy, ov = oracle.select(provided[0], p, 100, sort_kind=oracle_mod.SORT_STABLE)
        x, y, v = _get(capi, fl)
        assert np.array_equal(v, ov) and np.array_equal(x, ox) and np.array_equal(y, oy)
        L.track(tc, provided[0], provided[1], fl)
        gx, gy, gv = _get(capi, fl)
        tx, ty, tv = oracle.track(oracle.build_pyramids(provided[0], p), oracle.build_pyramids(provided[1], p),
                                  p, ox, oy, ov)
        assert np.array_equal(gv, tv) and gx.tobytes() == tx.tobytes() and gy.tobytes() == ty.tobytes()
        L.KLTFreeFeatureList(fl)
        L.KLTFreeTrackingContext(tc)


def test_device_frames_with_odd_pitch_and_offset(L, capi):
    """KLTTrackFeaturesDevice with a frame that is neither 16 B aligned nor 16 B pitched (no TMA
    descriptor possible): the tiled kernels take over, same result as the host-frame call."""
    import ctypes as C
    import torch
    h, w = 300, 403
    imgs = [synth_image(w, h, seed=31, shift=(1.3 * t, -0.6 * t)) for t in range(3)]
    res = []
    for mode in ("host", "device"):
        tc = L.KLTCreateTrackingContext()
        tc.contents.sequentialMode = 1
        fl = L.KLTCreateFeatureList(120)
        L.select(tc, imgs[0], fl)
        if mode == "host":
            for k in (1, 2):
                L.track(tc, imgs[k - 1], imgs[k], fl)
        else:
            pitch = w + 5
            bufs = []
            for im in imgs:
                buf = torch.zeros(h * pitch + 3, dtype=torch.uint8, device="cuda")
                view = buf[3:3 + h * pitch].view(h, pitch)
                view[:, :w] = torch.from_numpy(im).cuda()
                bufs.append(buf)
            torch.cuda.synchronize()
            for k in (1, 2):
                L.KLTTrackFeaturesDevice(tc, C.c_void_p(bufs[k - 1].data_ptr() + 3), C.c_void_p(bufs[k].data_ptr() + 3),
                                         pitch, w, h, fl)
        res.append(tuple(a.tobytes() for a in capi.featurelist_to_arrays(fl)))
        L.KLTFreeFeatureList(fl)
        L.KLTFreeTrackingContext(tc)
    assert res[0] == res[1]


def test_two_devices_in_one_process(L, capi, provided):
    """one process, tracking contexts on two GPUs (function attributes such as the dynamic
    shared-memory limit are per device): both give the same lists.  Needs >= 2 GPUs."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    got = []
    for dev in (0, 1, 0):
        tc = L.KLTCreateTrackingContext()
        tc.contents.sequentialMode = 1
        L.KLTB200SetDevice(tc, dev)
        fl = L.KLTCreateFeatureList(150)
        L.select(tc, provided[0], fl)
        for i in (1, 2):
            L.track(tc, provided[i - 1], provided[i], fl)
        L.replace(tc, provided[2], fl)
        got.append(_get(capi, fl))
        L.KLTFreeFeatureList(fl)
        L.KLTFreeTrackingContext(tc)
    for a in got[1:]:
        for u, v in zip(got[0], a):
            assert u.tobytes() == v.tobytes()


@pytest.mark.parametrize("switch", ["KLT_B200_NO_FILTER", "KLT_B200_MINEIG_SCALAR", "KLT_B200_REPLACE_FILTER"])
def test_selection_alternatives_are_bit_identical(L, capi, switch, monkeypatch):
    """the A/B switches of the selection path (whole-list walk in one launch; one thread per
    eigenvalue candidate) give the lists of the default path: select on an empty list (the walk's
    head / filter / tail hand-over at 640x480: 255 744 candidates), then replacement after losses."""
    w, h, n = 640, 480, 1200
    img0, img1 = synth_image(w, h, 31), synth_image(w, h, 31, shift=(1.5, -0.75))
    got = []
    for on in (0, 1):
        if on:
            monkeypatch.setenv(switch, "1")      # read when the context first selects
        tc = L.KLTCreateTrackingContext()
        tc.contents.sequentialMode = 1
        fl = L.KLTCreateFeatureList(n)
        L.select(tc, img0, fl)
        a = [u.copy() for u in _get(capi, fl)]
        L.track(tc, img0, img1, fl)
        x, y, v = [u.copy() for u in _get(capi, fl)]
        lost = np.arange(n) % 7 == 0
        x[lost], y[lost], v[lost] = -1.0, -1.0, -1
        capi.arrays_to_featurelist(fl, x, y, v)
        L.replace(tc, img1, fl)
        got.append(a + [u.copy() for u in _get(capi, fl)])
        L.KLTFreeFeatureList(fl)
        L.KLTFreeTrackingContext(tc)
    assert (got[0][2] > 0).sum() > 500
    for u, v in zip(got[0], got[1]):
        assert u.tobytes() == v.tobytes()
