"""BASELINE.json configurations at their real sizes on the GPU.

config 2 / 3 : real 640x480 frames (traffic, laptops), 1000 / 2000 features, with
               KLTReplaceLostFeatures every frame for config 3 -- vs the oracle.
config 4     : synthetic 3840x2160, 4096 features, 4 levels, subsampling 2 -- exact mode
               bit-identical to the oracle (images, selection, tracking), fma mode within
               tolerance, plus size-independent properties (translation recovered,
               determinism, resident pipeline == synchronous API).
config 5     : 1920x1080, 1024 features, default pyramid -- teacher-forced vs the oracle.
"""
import ctypes as C
import os

import numpy as np
import pytest

from tests.gpu_common import REL_TOL_IMAGES, check_fma_step, params_from_tc, rel_err

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def L(pkg):
    from importlib import import_module
    lib = import_module(pkg.__name__ + ".runtime").load()
    lib.require_gpu()
    lib.KLTSetVerbosity(0)
    return lib


@pytest.fixture(scope="module")
def synth(pkg):
    from importlib import import_module
    return import_module(pkg.__name__ + ".synth")


@pytest.fixture(scope="module")
def frames4k(synth):
    return [synth.frame(3840, 2160, seed=12345, t=float(t)) for t in range(3)]


def _tc4k(L, exact):
    tc = L.KLTCreateTrackingContext()
    tc.contents.sequentialMode = 1
    tc.contents.nPyramidLevels, tc.contents.subsampling = 4, 2
    L.KLTUpdateTCBorder(tc)
    L.KLTB200SetExact(tc, exact)
    return tc


def test_config4_images_exact_and_fma(L, oracle, frames4k):
    tc = _tc4k(L, 1)
    dev = L.KLTB200Device(tc)
    p = params_from_tc(oracle, tc)
    want = oracle.build_pyramids(frames4k[0], p)
    for exact in (1, 0):
        q = L.build_desc(tc, 3840, 2160, exact=exact)
        L.dev_build(dev, 0, frames4k[0], q)
        assert L.klt_dev_last_build_fused(dev) == 4
        for which in range(3):
            for l in range(4):
                a, b = L.dev_level(dev, 0, which, l), want.level(which, l)
                if exact:
                    assert np.array_equal(a, b), (which, l)
                else:
                    assert rel_err(a, b).max() <= REL_TOL_IMAGES
    L.KLTFreeTrackingContext(tc)


def test_config4_select_and_track_vs_oracle(L, capi, oracle, oracle_mod, frames4k):
    n = 4096
    p = None
    sel = None
    for exact in (1, 0):
        tc = _tc4k(L, exact)
        p = params_from_tc(oracle, tc)
        fl = L.KLTCreateFeatureList(n)
        L.select(tc, frames4k[0], fl)
        x, y, v = capi.featurelist_to_arrays(fl)
        if sel is None:
            sel = oracle.select(frames4k[0], p, n, sort_kind=oracle_mod.SORT_STABLE)
            pyr = [oracle.build_pyramids(f, p) for f in frames4k]
        assert np.array_equal(v, sel[2]) and np.array_equal(x, sel[0]) and np.array_equal(y, sel[1])
        assert (v > 0).sum() == n
        ox, oy, ov = sel
        for i in (1, 2):
            capi.arrays_to_featurelist(fl, ox, oy, ov)          # teacher forcing
            L.track(tc, frames4k[i - 1], frames4k[i], fl)
            gx, gy, gv = capi.featurelist_to_arrays(fl)
            x0, y0, v0 = ox, oy, ov
            ox, oy, ov = oracle.track(pyr[i - 1], pyr[i], p, ox, oy, ov)
            if exact:
                assert gx.tobytes() == ox.tobytes() and gy.tobytes() == oy.tobytes()
                assert np.array_equal(gv, ov)
            else:
                check_fma_step(oracle, p, pyr[i - 1], pyr[i], x0, y0, v0, gx, gy, gv, ox, oy, ov, "4K frame %d" % i)
        # the texture moves by -(2.3, -1.4) px per frame
        ok = ov >= 0
        assert ok.mean() > 0.97
        assert abs(np.median(ox[ok] - sel[0][ok]) + 2 * 2.3) < 0.1
        assert abs(np.median(oy[ok] - sel[1][ok]) - 2 * 1.4) < 0.1
        L.KLTFreeFeatureList(fl)
        L.KLTFreeTrackingContext(tc)


def test_config4_determinism_resident_and_device_paths(L, capi, frames4k):
    """same inputs -> same bytes: twice through the synchronous API, through the device-frame
    entry point and through the resident pipeline (no oracle needed: size-independent)."""
    import torch
    n = 4096
    results = []
    d_frames = [torch.from_numpy(f).cuda() for f in frames4k]
    torch.cuda.synchronize()
    for mode in ("host", "host", "device", "resident"):
        tc = _tc4k(L, 0)
        fl = L.KLTCreateFeatureList(n)
        L.select(tc, frames4k[0], fl)
        if mode == "host":
            for i in (1, 2):
                L.track(tc, frames4k[i - 1], frames4k[i], fl)
        elif mode == "device":
            for i in (1, 2):
                L.KLTTrackFeaturesDevice(tc, C.c_void_p(d_frames[i - 1].data_ptr()),
                                         C.c_void_p(d_frames[i].data_ptr()), 3840, 3840, 2160, fl)
        else:
            L.KLTB200ResidentBegin(tc, C.c_void_p(d_frames[0].data_ptr()), 1, 3840, 3840, 2160, fl)
            for i in (1, 2):
                L.KLTB200ResidentStep(tc, C.c_void_p(d_frames[i].data_ptr()), 1, 3840, 3840, 2160)
            L.KLTB200ResidentEnd(tc, fl)
        results.append(tuple(a.tobytes() for a in capi.featurelist_to_arrays(fl)))
        L.KLTFreeFeatureList(fl)
        L.KLTFreeTrackingContext(tc)
    assert results[0] == results[1] == results[2] == results[3]


def test_config5_1080p_default_pyramid(L, capi, oracle, oracle_mod, synth):
    frames = [synth.frame(1920, 1080, seed=1003, t=float(t), velocity=(-2.1, 1.7)) for t in range(4)]
    n = 1024
    for exact in (1, 0):
        tc = L.KLTCreateTrackingContext()
        tc.contents.sequentialMode = 1
        L.KLTB200SetExact(tc, exact)
        p = params_from_tc(oracle, tc)
        fl = L.KLTCreateFeatureList(n)
        L.select(tc, frames[0], fl)
        ox, oy, ov = oracle.select(frames[0], p, n, sort_kind=oracle_mod.SORT_STABLE)
        x, y, v = capi.featurelist_to_arrays(fl)
        assert np.array_equal(v, ov) and np.array_equal(x, ox) and np.array_equal(y, oy)
        prev = oracle.build_pyramids(frames[0], p)
        for i in range(1, 4):
            capi.arrays_to_featurelist(fl, ox, oy, ov)
            L.track(tc, frames[i - 1], frames[i], fl)
            gx, gy, gv = capi.featurelist_to_arrays(fl)
            cur = oracle.build_pyramids(frames[i], p)
            x0, y0, v0 = ox, oy, ov
            ox, oy, ov = oracle.track(prev, cur, p, ox, oy, ov)
            if exact:
                assert gx.tobytes() == ox.tobytes() and gy.tobytes() == oy.tobytes() and np.array_equal(gv, ov)
            else:
                check_fma_step(oracle, p, prev, cur, x0, y0, v0, gx, gy, gv, ox, oy, ov, "1080p frame %d" % i)
            prev = cur
        L.KLTFreeFeatureList(fl)
        L.KLTFreeTrackingContext(tc)


@pytest.mark.parametrize("dataset,n,replace", [("images_traffic", 1000, False), ("images_laptops", 2000, True)])
def test_config2_config3_real_frames(L, capi, oracle, oracle_mod, dataset, n, replace):
    imgs = [capi.read_pgm_numpy(os.path.join(GOLDEN, dataset, "img%d.pgm" % i)) for i in (1, 2, 3, 4)]
    assert imgs[0].shape == (480, 640)
    for exact in (1, 0):
        tc = L.KLTCreateTrackingContext()
        tc.contents.sequentialMode = 1
        L.KLTB200SetExact(tc, exact)
        p = params_from_tc(oracle, tc)
        fl = L.KLTCreateFeatureList(n)
        L.select(tc, imgs[0], fl)
        ox, oy, ov = oracle.select(imgs[0], p, n, sort_kind=oracle_mod.SORT_STABLE)
        x, y, v = capi.featurelist_to_arrays(fl)
        assert np.array_equal(v, ov) and np.array_equal(x, ox) and np.array_equal(y, oy)
        prev = oracle.build_pyramids(imgs[0], p)
        for i in range(1, 4):
            capi.arrays_to_featurelist(fl, ox, oy, ov)
            L.track(tc, imgs[i - 1], imgs[i], fl)
            gx, gy, gv = capi.featurelist_to_arrays(fl)
            cur = oracle.build_pyramids(imgs[i], p)
            x0, y0, v0 = ox, oy, ov
            ox, oy, ov = oracle.track(prev, cur, p, ox, oy, ov)
            if exact:
                assert gx.tobytes() == ox.tobytes() and gy.tobytes() == oy.tobytes() and np.array_equal(gv, ov)
            else:
                check_fma_step(oracle, p, prev, cur, x0, y0, v0, gx, gy, gv, ox, oy, ov, "%s frame %d" % (dataset, i))
            if replace:
                # config 3: replacement on the level 0 of this frame (exact arithmetic in both modes:
                # klt_dev_exact_level0), from the oracle's tracked list -> bit-identical in both modes
                capi.arrays_to_featurelist(fl, ox, oy, ov)
                L.replace(tc, imgs[i], fl)
                ox, oy, ov = oracle.select(imgs[i], p, n, sort_kind=oracle_mod.SORT_STABLE,
                                           replace=True, last=cur, x=ox, y=oy, val=ov)
                rx, ry, rv = capi.featurelist_to_arrays(fl)
                assert np.array_equal(rv, ov) and np.array_equal(rx, ox) and np.array_equal(ry, oy)
            prev = cur
        L.KLTFreeFeatureList(fl)
        L.KLTFreeTrackingContext(tc)
