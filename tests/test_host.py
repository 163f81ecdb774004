"""Host-side logic of libklt_b200.so that needs no GPU: ABI layout, parameter
derivation, containers, file formats, and that the library exports every
symbol the headers declare.  (No compute entry point is called here.)"""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def L(pkg):
    from importlib import import_module
    lib = import_module(pkg.__name__ + ".runtime").load()
    lib.KLTSetVerbosity(0)
    return lib


def test_abi_struct_sizes(capi):
    # SURVEY appendix C.1 (offsetof probe on the reference's klt.h)
    assert C.sizeof(capi.KLT_FeatureRec) == 64
    assert C.sizeof(capi.KLT_TrackingContextRec) == 136
    assert C.sizeof(capi.KLT_FeatureListRec) == 16
    assert C.sizeof(capi.KLT_FeatureTableRec) == 16
    assert capi.KLT_TrackingContextRec.pyramid_last.offset == 112
    assert capi.KLT_TrackingContextRec.borderx.offset == 68
    assert capi.KLT_FeatureRec.aff_img.offset == 16
    assert capi.KLT_FeatureRec.aff_Ayy.offset == 60


def test_abi_matches_c_compiler():
    """the header itself, compiled by gcc, gives the same layout"""
    src = r'''
    #include <stdio.h>
    #include <stddef.h>
    #include "klt.h"
    int main(void) {
      printf("%zu %zu %zu %zu %zu %zu\n", sizeof(KLT_FeatureRec), sizeof(KLT_TrackingContextRec),
             offsetof(KLT_TrackingContextRec, pyramid_last), offsetof(KLT_TrackingContextRec, nPyramidLevels),
             offsetof(KLT_FeatureRec, aff_x), sizeof(KLT_FeatureTableRec));
      return 0; }'''
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "a.c")
        open(c, "w").write(src)
        exe = os.path.join(td, "a.out")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        out = subprocess.check_output([exe]).decode().split()
    assert out == ["64", "136", "112", "76", "40", "16"]


def _declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", txt)
    return sorted({n for n in names if n.startswith(("KLT", "klt_dev_", "pgm", "ppm", "_KLT"))})


def test_library_exports_every_declared_symbol(L):
    for hdr in ("klt.h", "pnmio.h", "klt_cuda.h", "klt_b200.h"):
        names = _declared(hdr)
        assert names, hdr
        for n in names:
            assert hasattr(L.lib, n), "%s declares %s but the library does not export it" % (hdr, n)
    assert len(_declared("klt.h")) >= 31          # the 29 prototypes + KLTError/KLTWarning
    assert len(_declared("pnmio.h")) == 6


def test_defaults_and_parameter_derivation(L, oracle):
    tc = L.KLTCreateTrackingContext()
    t = tc.contents
    assert (t.mindist, t.window_width, t.window_height) == (10, 7, 7)
    assert (t.nPyramidLevels, t.subsampling, t.borderx, t.bordery) == (2, 4, 24, 24)
    assert t.sequentialMode == 0 and t.smoothBeforeSelecting == 1 and t.affineConsistencyCheck == -1
    assert abs(t.min_determinant - 0.01) < 1e-9 and t.max_iterations == 10
    assert t.pyramid_last is None
    p = oracle.default_params()
    for w in (3, 5, 7, 9, 11, 15):
        for sr in (1, 2, 4, 7, 10, 15, 16, 24, 40, 70, 150, 600):
            t.window_width = t.window_height = w
            p.window_width = p.window_height = w
            L.KLTChangeTCPyramid(tc, sr)
            L.KLTUpdateTCBorder(tc)
            oracle.change_pyramid(p, sr)
            oracle.update_border(p)
            assert (t.nPyramidLevels, t.subsampling, t.borderx, t.bordery) == \
                (p.nPyramidLevels, p.subsampling, p.borderx, p.bordery), (w, sr)
    t.window_width = t.window_height = 7
    t.nPyramidLevels, t.subsampling = 4, 2
    L.KLTUpdateTCBorder(tc)
    assert t.borderx == 64
    # even / tiny windows are repaired (with a warning on stderr)
    t.window_width, t.window_height = 6, 1
    L.KLTUpdateTCBorder(tc)
    assert (t.window_width, t.window_height) == (7, 3)
    L.KLTFreeTrackingContext(tc)


def test_taps_match_oracle(L, oracle):
    tc = L.KLTCreateTrackingContext()
    for ww, gs, ss in ((7, 1.0, 4), (7, 1.0, 2), (11, 1.8, 8), (3, 0.6, 2)):
        t = tc.contents
        t.window_width = t.window_height = ww
        t.grad_sigma, t.subsampling = gs, ss
        q = L.build_desc(tc, 64, 64)
        for taps, sigma in ((q.smooth_taps, np.float32(0.1) * ww), (q.pyramid_taps, ss * np.float32(0.9)),
                            (q.grad_taps, gs)):
            g, d = oracle.taps(float(np.float32(sigma)))
            assert taps.gauss_width == len(g) and taps.deriv_width == len(d)
            assert np.array_equal(np.array(taps.gauss[:len(g)], np.float32), g)
            assert np.array_equal(np.array(taps.deriv[:len(d)], np.float32), d)
    L.KLTFreeTrackingContext(tc)


def test_feature_containers_and_store_extract(L, capi):
    fl = L.KLTCreateFeatureList(5)
    ft = L.KLTCreateFeatureTable(3, 5)
    fh = L.KLTCreateFeatureHistory(3)
    assert fl.contents.nFeatures == 5 and ft.contents.nFrames == 3 and ft.contents.nFeatures == 5
    # one malloc block: header, pointer array, records
    base = C.addressof(fl.contents)
    assert C.addressof(fl.contents.feature[0].contents) == base + 16 + 5 * 8
    assert fl.contents.feature[0].contents.aff_img is None
    x = np.arange(5, dtype=np.float32) + 0.25
    capi.arrays_to_featurelist(fl, x, x * 2, np.array([3, 0, -1, -4, 7]))
    assert L.KLTCountRemainingFeatures(fl) == 3
    L.KLTStoreFeatureList(fl, ft, 1)
    tab = capi.featuretable_to_array(ft)
    assert np.array_equal(tab["x"][:, 1], x) and np.array_equal(tab["val"][:, 1], [3, 0, -1, -4, 7])
    capi.arrays_to_featurelist(fl, x * 0, x * 0, np.zeros(5))
    L.KLTExtractFeatureList(fl, ft, 1)
    assert np.array_equal(capi.featurelist_to_arrays(fl)[1], x * 2)
    L.KLTExtractFeatureHistory(fh, ft, 4)
    assert fh.contents.feature[1].contents.val == 7
    fh.contents.feature[2].contents.x = 9.5
    L.KLTStoreFeatureHistory(fh, ft, 0)
    assert capi.featuretable_to_array(ft)["x"][0, 2] == 9.5
    L.KLTFreeFeatureHistory(fh)
    L.KLTFreeFeatureTable(ft)
    L.KLTFreeFeatureList(fl)


def _golden_table(L, capi, golden_ft):
    _, gold = golden_ft
    ft = L.KLTCreateFeatureTable(10, 150)
    for j in range(150):
        for i in range(10):
            r = ft.contents.feature[j][i].contents
            r.x, r.y, r.val = float(gold["x"][j, i]), float(gold["y"][j, i]), int(gold["val"][j, i])
    return ft


def test_writers_reproduce_golden_files(L, capi, golden_ft, tmp_path):
    raw, _ = golden_ft
    ft = _golden_table(L, capi, golden_ft)
    b = str(tmp_path / "t.ft").encode()
    t = str(tmp_path / "t.txt").encode()
    L.KLTWriteFeatureTable(ft, b, None)
    L.KLTWriteFeatureTable(ft, t, b"%5.1f")
    assert open(b, "rb").read() == raw
    assert open(t, "rb").read() == open(os.path.join(GOLDEN, "features2.txt"), "rb").read()
    # round trip through the readers, both formats
    for path, exact in ((b, True), (t, False)):
        back = L.KLTReadFeatureTable(None, path)
        arr = capi.featuretable_to_array(back)
        want = capi.featuretable_to_array(ft)
        assert np.array_equal(arr["val"], want["val"])
        if exact:
            assert arr.tobytes() == want.tobytes()
        else:
            assert np.abs(arr["x"] - want["x"]).max() <= 0.05 + 1e-6
        L.KLTFreeFeatureTable(back)
    L.KLTFreeFeatureTable(ft)


def test_list_and_history_files_round_trip(L, capi, tmp_path):
    fl = L.KLTCreateFeatureList(4)
    capi.arrays_to_featurelist(fl, [1.5, 2.25, -1, 300.75], [4, 5.5, -1, 0.5], [12, 0, -4, 3])
    for fmt in (None, b"%7.2f", b"%3d"):
        path = str(tmp_path / "l").encode()
        L.KLTWriteFeatureList(fl, path, fmt)
        back = L.KLTReadFeatureList(None, path)
        x, y, v = capi.featurelist_to_arrays(back)
        assert np.array_equal(v, [12, 0, -4, 3])
        if fmt != b"%3d":
            assert np.array_equal(x, np.array([1.5, 2.25, -1, 300.75], np.float32))
        else:
            assert np.array_equal(x, np.array([2, 2, -1, 301], np.float32))
        L.KLTFreeFeatureList(back)
    fh = L.KLTCreateFeatureHistory(3)
    for i in range(3):
        r = fh.contents.feature[i].contents
        r.x, r.y, r.val = i + 0.5, 2 * i, i - 1
    path = str(tmp_path / "h").encode()
    L.KLTWriteFeatureHistory(fh, path, b"%5.1f")
    back = L.KLTReadFeatureHistory(None, path)
    assert [back.contents.feature[i].contents.val for i in range(3)] == [-1, 0, 1]
    L.KLTFreeFeatureHistory(back)
    L.KLTFreeFeatureHistory(fh)
    L.KLTFreeFeatureList(fl)


def test_pnm_io_and_overlay(L, capi, provided, tmp_path):
    src = os.path.join(GOLDEN, "images_provided", "img0.pgm")
    img = L.read_pgm(src)                      # through the library's pgmReadFile ('#' comment in header)
    assert img.shape == (240, 320) and np.array_equal(img, provided[0])
    out = str(tmp_path / "o.pgm").encode()
    L.pgmWriteFile(out, img.ctypes.data_as(C.c_void_p), 320, 240)
    assert open(out, "rb").read() == b"P5\n320 240\n255\n" + img.tobytes()
    assert np.array_equal(capi.read_pgm_numpy(out.decode()), img)
    fl = L.KLTCreateFeatureList(3)
    capi.arrays_to_featurelist(fl, [10.4, 0.2, -1], [20.6, 0.0, -1], [5, 0, -4])
    ppm = str(tmp_path / "o.ppm").encode()
    L.KLTWriteFeatureListToPPM(fl, img.ctypes.data_as(C.c_void_p), 320, 240, ppm)
    data = open(ppm, "rb").read()
    hdr = b"P6\n320 240\n255\n"
    assert data.startswith(hdr)
    rgb = np.frombuffer(data[len(hdr):], np.uint8).reshape(240, 320, 3)
    want = np.repeat(img[:, :, None], 3, axis=2).copy()
    want[20:23, 9:12] = (255, 0, 0)            # round(10.4)=10, round(20.6)=21
    want[0:2, 0:2] = (255, 0, 0)               # clipped at the corner
    assert np.array_equal(rgb, want)
    L.KLTFreeFeatureList(fl)


def test_pgm_header_comment_placements(L, tmp_path):
    """'#' comments anywhere in the header, including one that abuts maxval ("255#c\\n": its newline
    IS the single separator, so no further byte may be skipped -- the reference, pnmio.c:66-69,
    loses the first pixel there)."""
    px = (np.arange(6 * 4, dtype=np.uint8) * 9 + 7).reshape(4, 6)
    for k, hdr in enumerate([b"P5\n6 4\n255\n", b"P5\n# made by x\n6 4\n255\n", b"P5 6 4 255\n",
                             b"P5\n6 4 # size\n255\n", b"P5\n6 4\n255#c\n", b"P5\n6 4\n255 # c\n"[:0] or b"P5\n6#w\n4\n255\n"]):
        f = tmp_path / ("h%d.pgm" % k)
        f.write_bytes(hdr + px.tobytes())
        got = L.read_pgm(str(f))
        assert got.shape == (4, 6) and np.array_equal(got, px), hdr


def test_hot_path_fails_loudly_without_gpu(L):
    """No CPU fallback: on a box without a CUDA device the hot path is a
    KLTError (message + exit(1)), never a silent CPU computation."""
    if L.device_count() > 0:
        pytest.skip("a GPU is present")
    code = ("import importlib,numpy as np;"
            "m=importlib.import_module('klt-feature-tracker-acceleration-gpus_b200.runtime');"
            "L=m.load();tc=L.KLTCreateTrackingContext();fl=L.KLTCreateFeatureList(4);"
            "L.select(tc,np.zeros((64,64),np.uint8),fl);print('SURVIVED')")
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True)
    assert r.returncode == 1
    assert "KLT Error" in r.stderr and "SURVIVED" not in r.stdout
    with pytest.raises(RuntimeError):
        L.require_gpu()


def test_contexts_from_many_threads(L):
    """The side table tc -> device state behind a per-thread one-entry cache (csrc/klt_context.c): four
    threads create, configure and free contexts in a loop -- every free voids the other threads' cached
    entries -- and a context freed by one thread and re-created at the same address by another must not
    find the old state (KLTB200SetExact is stored in the state: a fresh context starts with 0)."""
    import threading
    errors = []

    def worker(seed):
        try:
            for it in range(300):
                tcs = [L.KLTCreateTrackingContext() for _ in range(3)]
                for k, tc in enumerate(tcs):
                    assert L.KLTB200GetExact(tc) == 0, "stale state on a fresh context"
                    L.KLTB200SetExact(tc, (seed + k + it) & 1)
                for k, tc in enumerate(tcs):
                    assert L.KLTB200GetExact(tc) == ((seed + k + it) & 1)
                for tc in tcs:
                    L.KLTFreeTrackingContext(tc)
        except Exception as e:                       # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[:3]
