"""Constructed status-code edge cases of _trackFeature (reference
src/V1/trackFeatures.c:409-424 the 1.001 out-of-bounds slack, :293-307 det < min_determinant ->
KLT_SMALL_DET, :459-474 residue > max_residue -> KLT_LARGE_RESIDUE, :457/:481-482
KLT_MAX_ITERATIONS, :1346 lost features are skipped).

Real frames reach these branches only by accident (the survey counted 0 SMALL_DET events in all of
config 2), so each case here is built to sit ON the threshold: populations of features that straddle
it from both sides.  Exact mode must reproduce the oracle bit for bit -- positions and status codes
-- and the default fma mode must agree within the north_star tolerance with every disagreement
explained by threshold proximity (tests/gpu_common.check_fma_step).  Each test also asserts that
both sides of its threshold are actually populated, so that it cannot pass vacuously.
"""
import numpy as np
import pytest

from tests.conftest import synth_image
from tests.gpu_common import check_fma_step, params_from_tc

pytestmark = pytest.mark.gpu

KLT_TRACKED, KLT_NOT_FOUND, KLT_SMALL_DET, KLT_MAX_ITERATIONS, KLT_OOB, KLT_LARGE_RESIDUE = 0, -1, -2, -3, -4, -5


@pytest.fixture(scope="module")
def L(pkg):
    from importlib import import_module
    lib = import_module(pkg.__name__ + ".runtime").load()
    lib.require_gpu()
    lib.KLTSetVerbosity(0)
    return lib


def _run_pair(L, capi, oracle, img1, img2, x, y, exact, setup=None):
    """one KLTTrackFeatures call on hand-placed features -> (gpu result, oracle result, params, pyramids)"""
    n = len(x)
    tc = L.KLTCreateTrackingContext()
    if setup:
        setup(tc)
    L.KLTB200SetExact(tc, exact)
    p = params_from_tc(oracle, tc)
    fl = L.KLTCreateFeatureList(n)
    v = np.zeros(n, np.int32)
    capi.arrays_to_featurelist(fl, x, y, v)
    L.track(tc, img1, img2, fl)
    g = capi.featurelist_to_arrays(fl)
    p1, p2 = oracle.build_pyramids(img1, p), oracle.build_pyramids(img2, p)
    o = oracle.track(p1, p2, p, x, y, v)
    L.KLTFreeFeatureList(fl)
    L.KLTFreeTrackingContext(tc)
    return g, o, p, p1, p2


def _compare(oracle, g, o, p, p1, p2, x, y, exact, where):
    gx, gy, gv = g
    ox, oy, ov = o
    if exact:
        assert np.array_equal(gv, ov), (where, np.nonzero(gv != ov)[0][:8], gv[gv != ov][:8], ov[gv != ov][:8])
        assert gx.tobytes() == ox.tobytes() and gy.tobytes() == oy.tobytes(), where
        return 0
    return check_fma_step(oracle, p, p1, p2, np.asarray(x, np.float32), np.asarray(y, np.float32),
                          np.zeros(len(x), np.int32), gx, gy, gv, ox, oy, ov, where, fraction_gate=False)


@pytest.mark.parametrize("levels", [2, 1])
@pytest.mark.parametrize("exact", [1, 0])
def test_small_determinant_flat_and_faint_patches(L, capi, oracle, exact, levels):
    """det = gxx*gyy - gxy^2 < 0.01 -> KLT_SMALL_DET (:293-307).  The frame is flat (128) except
    for two rows of texture patches whose amplitude rises from nothing to clearly trackable;
    features sit at the patch centres, so det sweeps through min_determinant.  Two more features sit
    on flat ground near the left edge: with two pyramid levels their KLT_SMALL_DET is raised at the
    coarse level, where the descent stops (:1378) with the output position still in coarse
    coordinates -- inside the border -- so the reference records KLT_OOB for them (:1393); with one
    level they are plain KLT_SMALL_DET."""
    W, H = 640, 320
    tex = synth_image(W, H, seed=77).astype(np.float32) - 128.0
    img = np.full((H, W), 128.0, np.float32)
    xs, ys = [], []
    amps = np.concatenate([np.zeros(4), np.geomspace(0.002, 0.6, 28)])
    k = 0
    for row in range(2):
        for col in range(16):
            cx, cy = 112 + col * 32, 120 + row * 120
            a = amps[k]; k += 1
            img[cy - 15:cy + 16, cx - 15:cx + 16] += a * tex[cy - 15:cy + 16, cx - 15:cx + 16]
            xs.append(cx + 0.25); ys.append(cy - 0.5)
    xs += [40.0, 44.5]; ys += [120.0, 240.0]
    img1 = np.clip(np.rint(img), 0, 255).astype(np.uint8)
    img2 = np.roll(img1, 1, axis=1)                    # 1 px to the right
    x, y = np.array(xs, np.float32), np.array(ys, np.float32)

    def setup(tc):
        tc.contents.nPyramidLevels = levels            # (borders stay at the default 24)

    g, o, p, p1, p2 = _run_pair(L, capi, oracle, img1, img2, x, y, exact, setup)
    ov = o[2]
    assert (ov == KLT_SMALL_DET).sum() >= 6 and (ov == KLT_TRACKED).sum() >= 6, ov
    assert (ov[:4] == KLT_SMALL_DET).all()             # the perfectly flat patches
    assert (ov[-2:] == (KLT_OOB if levels == 2 else KLT_SMALL_DET)).all(), ov[-2:]
    _compare(oracle, g, o, p, p1, p2, x, y, exact, "small-det sweep")


@pytest.mark.parametrize("exact", [1, 0])
def test_out_of_bounds_slack_1001(L, capi, oracle, exact):
    """nc - (x + hw) < 1.001 (and the other three sides) -> KLT_OOB before anything is sampled
    (:409-424).  One pyramid level and a zero border, so that this test -- not the coarse level,
    not the final border check -- decides: features at nc-(x+hw) = 1.0005 must be lost, at 1.0015
    kept; x - hw = -0.0005 lost, +0.0005 kept; same in y."""
    W, H = 320, 240
    img1 = synth_image(W, H, seed=5)
    img2 = img1.copy()                                 # no motion: kept features converge at once
    hw = 3
    xs, ys = [], []
    for eps in (0.0005, 0.0009, 0.00101, 0.0011, 0.0015, 0.01, 0.5):
        for yy in (50.0, 120.25, 200.5):
            xs.append(W - hw - 1.0 - eps); ys.append(yy)          # right edge
            xs.append(hw - 0.001 + eps); ys.append(yy)            # left edge: x - hw < 0
        for xx in (60.0, 160.75, 250.5):
            xs.append(xx); ys.append(H - hw - 1.0 - eps)          # bottom edge
            xs.append(xx); ys.append(hw - 0.001 + eps)            # top edge
    x, y = np.array(xs, np.float32), np.array(ys, np.float32)

    def setup(tc):
        tc.contents.nPyramidLevels = 1
        tc.contents.borderx = tc.contents.bordery = 0

    g, o, p, p1, p2 = _run_pair(L, capi, oracle, img1, img2, x, y, exact, setup)
    ov = o[2]
    assert (ov == KLT_OOB).sum() >= 10 and (ov == KLT_TRACKED).sum() >= 10, np.bincount(-ov)
    # the decision is a comparison on the inputs: both arithmetic modes must make it identically
    assert np.array_equal(g[2] == KLT_OOB, ov == KLT_OOB), np.nonzero((g[2] == KLT_OOB) != (ov == KLT_OOB))[0]
    _compare(oracle, g, o, p, p1, p2, x, y, exact, "oob slack")


@pytest.mark.parametrize("levels,motion", [(1, 3.2), (2, 13.0)])
@pytest.mark.parametrize("exact", [1, 0])
def test_out_of_bounds_after_motion(L, capi, oracle, exact, levels, motion):
    """the test on (x2, y2) inside and after the Newton loop (:421-424, :459-461), at level 0 (one
    level, 3.2 px of motion) and at the coarse level of a two-level pyramid (13 px of motion; the
    descent stops there, :1378): the texture moves towards the right edge and carries the features
    nearest to it out of bounds, the others end in every other status."""
    W, H = 320, 240
    img1 = synth_image(W, H, seed=9)
    img2 = synth_image(W, H, seed=9, shift=(-motion, 0.0))        # content moves +motion px in x
    if levels == 1:
        xs = [W - 3 - 1.2 - 0.12 * k for k in range(40)]
    else:
        xs = [306 - 0.5 * k for k in range(40)]
    ys = [30.0 + 4.5 * k for k in range(40)]
    x, y = np.array(xs, np.float32), np.array(ys, np.float32)

    def setup(tc):
        tc.contents.nPyramidLevels = levels
        tc.contents.borderx = tc.contents.bordery = 0

    g, o, p, p1, p2 = _run_pair(L, capi, oracle, img1, img2, x, y, exact, setup)
    ov = o[2]
    assert (ov == KLT_OOB).sum() >= 5 and (ov == KLT_TRACKED).sum() >= 5, ov
    _compare(oracle, g, o, p, p1, p2, x, y, exact, "oob after motion")


@pytest.mark.parametrize("exact", [1, 0])
def test_residue_straddles_max_residue(L, capi, oracle, exact):
    """sum |I1 - I2| / (ww*wh) > max_residue -> KLT_LARGE_RESIDUE (:463-474).  Frame 2 is frame 1
    plus a second, unrelated texture whose weight rises from 0 at the left edge to 0.8 at the right:
    features converge (nearly) in place and their residue grows with x, through max_residue.  The
    threshold is swept as well, so that the population splits at five different places."""
    W, H = 640, 480
    img1 = synth_image(W, H, seed=31)
    noise = synth_image(W, H, seed=32).astype(np.float32) - 128.0
    a = np.linspace(0.0, 0.8, W, dtype=np.float32)[None, :]
    img2 = np.clip(np.rint(img1.astype(np.float32) + a * noise), 0, 255).astype(np.uint8)
    tc = L.KLTCreateTrackingContext()
    p0 = params_from_tc(oracle, tc)
    L.KLTFreeTrackingContext(tc)
    x, y, v = oracle.select(img1, p0, 600)
    x, y = x[v > 0], y[v > 0]
    for mr in (4.0, 6.0, 8.0, 10.0, 12.0):
        def setup(tc, mr=mr):
            tc.contents.max_residue = mr
        g, o, p, p1, p2 = _run_pair(L, capi, oracle, img1, img2, x, y, exact, setup)
        ov = o[2]
        assert (ov == KLT_LARGE_RESIDUE).sum() >= 15 and (ov == KLT_TRACKED).sum() >= 200, (mr, np.bincount(-ov))
        _compare(oracle, g, o, p, p1, p2, x, y, exact, "max_residue %g" % mr)


@pytest.mark.parametrize("max_it", [1, 2, 3])
@pytest.mark.parametrize("exact", [1, 0])
def test_max_iterations(L, capi, oracle, exact, max_it):
    """iteration >= max_iterations -> KLT_MAX_ITERATIONS (:457, :481-482), which does NOT stop the
    coarse-to-fine descent (:1378).  With 1..3 iterations allowed and 2.3 px of motion, part of
    the features run out of iterations at level 0 and part converge in time."""
    W, H = 640, 480
    img1 = synth_image(W, H, seed=41)
    img2 = synth_image(W, H, seed=41, shift=(2.3, -1.4))
    tc = L.KLTCreateTrackingContext()
    p0 = params_from_tc(oracle, tc)
    L.KLTFreeTrackingContext(tc)
    x, y, v = oracle.select(img1, p0, 500)
    x, y = x[v > 0], y[v > 0]

    def setup(tc):
        tc.contents.max_iterations = max_it

    g, o, p, p1, p2 = _run_pair(L, capi, oracle, img1, img2, x, y, exact, setup)
    ov = o[2]
    assert (ov == KLT_MAX_ITERATIONS).sum() >= 20, np.bincount(-ov)
    if max_it > 1:
        assert (ov == KLT_TRACKED).sum() >= 20, np.bincount(-ov)
    else:
        assert (ov == KLT_TRACKED).sum() == 0                      # one iteration always "runs out"
    _compare(oracle, g, o, p, p1, p2, x, y, exact, "max_iterations %d" % max_it)


@pytest.mark.parametrize("exact", [1, 0])
def test_lost_features_are_skipped_and_keep_their_code(L, capi, oracle, exact):
    """features with val < 0 are not tracked and not rewritten (:1346): whatever code and
    coordinates they carry come back unchanged, in both list layouts of the host API."""
    W, H = 320, 240
    img1 = synth_image(W, H, seed=3)
    img2 = synth_image(W, H, seed=3, shift=(1.1, 0.6))
    tc = L.KLTCreateTrackingContext()
    L.KLTB200SetExact(tc, exact)
    p = params_from_tc(oracle, tc)
    n = 120
    x, y, v = oracle.select(img1, p, n)
    v[:] = 0
    codes = [KLT_NOT_FOUND, KLT_SMALL_DET, KLT_MAX_ITERATIONS, KLT_OOB, KLT_LARGE_RESIDUE]
    for k in range(0, n, 3):
        v[k] = codes[(k // 3) % 5]
        x[k], y[k] = -1.0, -1.0
    x[6], y[6] = 33.5, 44.25                           # a lost feature with stale coordinates stays as it is
    fl = L.KLTCreateFeatureList(n)
    capi.arrays_to_featurelist(fl, x, y, v)
    L.track(tc, img1, img2, fl)
    gx, gy, gv = capi.featurelist_to_arrays(fl)
    ox, oy, ov = oracle.track(oracle.build_pyramids(img1, p), oracle.build_pyramids(img2, p), p, x, y, v)
    lost = v < 0
    assert np.array_equal(gv[lost], v[lost]) and np.array_equal(gx[lost], x[lost]) and np.array_equal(gy[lost], y[lost])
    assert np.array_equal(ov[lost], v[lost])
    if exact:
        assert np.array_equal(gv, ov) and gx.tobytes() == ox.tobytes() and gy.tobytes() == oy.tobytes()
    else:
        assert (gv == ov).mean() >= 0.99
    L.KLTFreeFeatureList(fl)
    L.KLTFreeTrackingContext(tc)
